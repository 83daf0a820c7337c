"""GECKO fragment records and the GECKO CSV container.

The in-memory record is the reference's ``struct FragFile`` under ``#pragma pack(1)``
(/root/reference/src/structs.h:2,12-51): 109 bytes, no padding.  Offsets:
diag i64@0, xStart u64@8, yStart u64@16, xEnd u64@24, yEnd u64@32, length u64@40, ident u64@48,
score u64@56, similarity f32@64, seqX u64@68, seqY u64@76, block i64@84, strand char@92,
evalue long double@93 (16 opaque bytes).

GECKO's binary container (``.frags``; csrc/host/GeckoFrags.h, SURVEY.md §8f N4 — not read by the reference, restated from
GECKO's published writer): two big-endian uint64 sequence lengths, then the same 109-byte records with every field stored
most significant byte first.

The CSV container is what the reference's ``FragmentsDatabase`` constructor reads
(/root/reference/src/FragmentsDatabase.cpp:54-101): 16 header lines (line 7 ``SeqX length``, line 8
``SeqY length``, line 13 ``Total fragments``), then one ``Frag,...`` row per fragment.
"""
from __future__ import annotations

import io
import numpy as np

FRAG_BYTES = 109

FRAG_DTYPE = np.dtype(
    {
        "names": ["diag", "xStart", "yStart", "xEnd", "yEnd", "length", "ident", "score",
                  "similarity", "seqX", "seqY", "block", "strand", "evalue"],
        "formats": ["<i8", "<u8", "<u8", "<u8", "<u8", "<u8", "<u8", "<u8",
                    "<f4", "<u8", "<u8", "<i8", "S1", "V16"],
        "offsets": [0, 8, 16, 24, 32, 40, 48, 56, 64, 68, 76, 84, 92, 93],
        "itemsize": FRAG_BYTES,
    }
)
assert FRAG_DTYPE.itemsize == FRAG_BYTES

HEADER_TEMPLATE = (
    "All by-Identity Ungapped Fragments (Hits based approach)\n"
    "[Abr.2015 -- < bitlab - Departamento de Arquitectura de Computadores >\n"
    "SeqX filename : synthX.fasta\n"
    "SeqY filename : synthY.fasta\n"
    "SeqX name : synthX\n"
    "SeqY name : synthY\n"
    "SeqX length : {lx}\n"
    "SeqY length : {ly}\n"
    "Min.fragment.length : 0\n"
    "Min.Identity : 0.00\n"
    "Tot Hits (seeds) : 0\n"
    "Tot Hits (seeds) used: 0\n"
    "Total fragments : {n}\n"
    "========================================================\n"
    "Type,xStart,yStart,xEnd,yEnd,strand(f/r),block,length,score,ident,similarity,%ident,SeqX,SeqY\n"
    "========================================================\n"
)


def make_header(lx_header: int, ly_header: int, total_frags: int) -> str:
    """16-line GECKO CSV header.  ``lx_header``/``ly_header`` are the values printed in the file;
    the reference adds 1 to each when loading (FragmentsDatabase.cpp:62,65)."""
    h = HEADER_TEMPLATE.format(lx=lx_header, ly=ly_header, n=total_frags)
    assert h.count("\n") == 16
    return h


def empty_records(n: int) -> np.ndarray:
    return np.zeros(n, dtype=FRAG_DTYPE)


def records_to_csv_rows(rec: np.ndarray) -> str:
    """``Frag,xStart,yStart,xEnd,yEnd,strand,block,length,score,ident,similarity,%ident,0,0`` rows.
    ``similarity`` is printed with %.9g so that the reference's ``stof`` recovers the same float."""
    out = io.StringIO()
    xs, ys, xe, ye = rec["xStart"], rec["yStart"], rec["xEnd"], rec["yEnd"]
    st, bl, ln, sc, idn, sim = rec["strand"], rec["block"], rec["length"], rec["score"], rec["ident"], rec["similarity"]
    for i in range(rec.shape[0]):
        s = "%.9g" % float(sim[i])
        out.write("Frag,%d,%d,%d,%d,%s,%d,%d,%d,%d,%s,%s,0,0\n" % (
            xs[i], ys[i], xe[i], ye[i], st[i].decode("latin1"), bl[i], ln[i], sc[i], idn[i], s, s))
    return out.getvalue()


def write_csv(path: str, rec: np.ndarray, lx_header: int, ly_header: int, total_frags: int | None = None) -> None:
    with open(path, "w", encoding="latin1", newline="") as f:
        f.write(make_header(lx_header, ly_header, rec.shape[0] if total_frags is None else total_frags))
        f.write(records_to_csv_rows(rec))


# ---- GECKO's binary container (.frags) ----------------------------------------------------------------------------------
FRAG_DTYPE_BE = np.dtype(
    {
        "names": FRAG_DTYPE.names,
        "formats": [">i8", ">u8", ">u8", ">u8", ">u8", ">u8", ">u8", ">u8", ">f4", ">u8", ">u8", ">i8", "S1", "V16"],
        "offsets": [0, 8, 16, 24, 32, 40, 48, 56, 64, 68, 76, 84, 92, 93],
        "itemsize": FRAG_BYTES,
    }
)
GECKO_HEADER_BYTES = 16


def records_to_gecko_binary(rec: np.ndarray, seqx_len: int, seqy_len: int) -> bytes:
    """header (two big-endian uint64) + every record with its fields byte-reversed (``evalue`` included)"""
    be = np.zeros(rec.shape[0], dtype=FRAG_DTYPE_BE)
    for name in FRAG_DTYPE.names:
        if name != "evalue":
            be[name] = rec[name]
    ev = np.ascontiguousarray(rec["evalue"]).view(np.uint8).reshape(-1, 16)[:, ::-1]
    be["evalue"] = np.ascontiguousarray(ev).view("V16").reshape(-1)
    return int(seqx_len).to_bytes(8, "big") + int(seqy_len).to_bytes(8, "big") + be.tobytes()


def write_gecko_binary(path: str, rec: np.ndarray, seqx_len: int, seqy_len: int) -> None:
    with open(path, "wb") as f:
        f.write(records_to_gecko_binary(rec, seqx_len, seqy_len))


def gecko_binary_as_loaded(data: bytes) -> tuple[np.ndarray, int, int]:
    """What a .frags file loads as (GeckoFrags.h): the record ``readFragment`` builds from a CSV row printing the stored values
    (/root/reference/src/FragmentsDatabase.cpp:30-43) — ident := trunc(similarity), diag := xStart - yStart, seqX 0, seqY 1,
    evalue 0 — and the two sequence lengths of the header."""
    if len(data) < GECKO_HEADER_BYTES or (len(data) - GECKO_HEADER_BYTES) % FRAG_BYTES:
        raise ValueError("not a GECKO binary fragments file")
    lx, ly = int.from_bytes(data[:8], "big"), int.from_bytes(data[8:16], "big")
    be = np.frombuffer(data, dtype=FRAG_DTYPE_BE, offset=GECKO_HEADER_BYTES)
    rec = empty_records(be.shape[0])
    for name in ("xStart", "yStart", "xEnd", "yEnd", "length", "score", "similarity", "block", "strand"):
        rec[name] = be[name]
    rec["diag"] = (rec["xStart"].astype(np.int64) - rec["yStart"].astype(np.int64))
    rec["ident"] = rec["similarity"].astype(np.uint64)   # (finite, non-negative similarities: what GECKO stores)
    rec["seqX"], rec["seqY"] = 0, 1
    return rec, lx, ly
