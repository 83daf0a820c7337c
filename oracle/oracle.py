"""TEST INFRASTRUCTURE — NOT part of the product path.

ctypes front end of the C restatement (oracle/rk_oracle.c) and a runner for the compiled reference
(oracle/_ref/repkiller_ref).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "librk_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "repkiller_ref")
NONE = 0xFFFFFFFF


def build(ref: bool = True) -> None:
    """Compile the C restatement and, when /root/reference is present, the reference behind ref_driver.cpp."""
    subprocess.check_call(["make", "-s", "-C", HERE, "lib"] + (["ref", "refbody"] if ref else []))


class _Result(C.Structure):
    _fields_ = [
        ("n_kept", C.c_uint64), ("n_groups", C.c_uint64), ("vsize", C.c_uint64),
        ("rank_fidx", C.POINTER(C.c_uint32)), ("xowner", C.POINTER(C.c_uint32)), ("yowner", C.POINTER(C.c_uint32)),
        ("parent", C.POINTER(C.c_uint32)), ("gid", C.POINTER(C.c_uint32)), ("h", C.POINTER(C.c_uint64)),
        ("order", C.POINTER(C.c_uint32)), ("out_gid", C.POINTER(C.c_uint32)), ("repval", C.POINTER(C.c_uint8)),
        ("identity", C.POINTER(C.c_float)), ("diag_func", C.POINTER(C.c_uint64)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = C.CDLL(LIB_PATH)
        L.rko_parse_row.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.rko_parse_row.restype = C.c_int
        L.rko_load_csv.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
        L.rko_load_csv.restype = C.c_int
        L.rko_group.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_int,
                                C.POINTER(_Result)]
        L.rko_group.restype = C.c_int
        L.rko_result_free.argtypes = [C.POINTER(_Result)]
        L.rko_write_output.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.POINTER(_Result)]
        L.rko_write_output.restype = C.c_int
        L.rko_write_input_csv.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.rko_write_input_csv.restype = C.c_int
        L.rko_free.argtypes = [C.c_void_p]
        L.rko_std_sort_by_key.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


@dataclass
class OracleGroups:
    n_kept: int
    n_groups: int
    vsize: int
    rank_fidx: np.ndarray
    xowner: np.ndarray
    yowner: np.ndarray
    parent: np.ndarray
    gid: np.ndarray
    h: np.ndarray
    order: np.ndarray
    out_gid: np.ndarray
    repval: np.ndarray
    identity: np.ndarray
    diag_func: np.ndarray | None


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _frag_dtype():
    from repkiller_b200.frags import FRAG_DTYPE
    return FRAG_DTYPE


def load_csv(path: str):
    """(records, lx1, ly1, header) the way the reference's FragmentsDatabase loads them (lengths are header + 1)."""
    L = lib()
    recs, hdr = C.c_void_p(), C.c_void_p()
    n, lx1, ly1 = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = L.rko_load_csv(path.encode(), C.byref(recs), C.byref(n), C.byref(lx1), C.byref(ly1), C.byref(hdr))
    if rc == -1:
        raise FileNotFoundError(path)
    dt = _frag_dtype()
    out = np.frombuffer(C.string_at(recs.value, n.value * dt.itemsize), dtype=dt).copy() if n.value else np.zeros(0, dt)
    header = C.string_at(hdr.value)
    L.rko_free(recs)
    L.rko_free(hdr)
    if rc == -2:
        raise RuntimeError("Unexpected number of fragments")
    return out, lx1.value, ly1.value, header


def group(records: np.ndarray, lx1: int, ly1: int, len_ratio: float, pos_ratio: float, want_diag: bool = False) -> OracleGroups:
    L = lib()
    rec = np.ascontiguousarray(records)
    assert rec.dtype.itemsize == 109
    r = _Result()
    rc = L.rko_group(rec.ctypes.data, rec.shape[0], lx1, ly1, len_ratio, pos_ratio, int(want_diag), C.byref(r))
    if rc:
        raise ValueError(f"rko_group failed: {rc}")
    m = r.n_kept
    out = OracleGroups(
        n_kept=m, n_groups=r.n_groups, vsize=r.vsize,
        rank_fidx=_arr(r.rank_fidx, m, np.uint32), xowner=_arr(r.xowner, m, np.uint32),
        yowner=_arr(r.yowner, m, np.uint32), parent=_arr(r.parent, m, np.uint32), gid=_arr(r.gid, m, np.uint32),
        h=_arr(r.h, m, np.uint64), order=_arr(r.order, m, np.uint32), out_gid=_arr(r.out_gid, m, np.uint32),
        repval=_arr(r.repval, m, np.uint8), identity=_arr(r.identity, m, np.float32),
        diag_func=_arr(r.diag_func, r.vsize - 1, np.uint64) if want_diag else None,
    )
    L.rko_result_free(C.byref(r))
    return out


def write_output(path: str, header: bytes, records: np.ndarray, g: OracleGroups) -> None:
    L = lib()
    r = _Result()
    r.n_kept = g.n_kept
    keep = [np.ascontiguousarray(g.order), np.ascontiguousarray(g.out_gid), np.ascontiguousarray(g.repval),
            np.ascontiguousarray(g.identity)]
    r.order = keep[0].ctypes.data_as(C.POINTER(C.c_uint32))
    r.out_gid = keep[1].ctypes.data_as(C.POINTER(C.c_uint32))
    r.repval = keep[2].ctypes.data_as(C.POINTER(C.c_uint8))
    r.identity = keep[3].ctypes.data_as(C.POINTER(C.c_float))
    rec = np.ascontiguousarray(records)
    if L.rko_write_output(path.encode(), header, rec.ctypes.data, C.byref(r)):
        raise OSError(f"cannot write {path}")


def write_input_csv(path: str, records: np.ndarray, lx_header: int, ly_header: int) -> None:
    rec = np.ascontiguousarray(records)
    if lib().rko_write_input_csv(path.encode(), rec.ctypes.data, rec.shape[0], lx_header, ly_header):
        raise OSError(f"cannot write {path}")


def group_statistics(records: np.ndarray, g: OracleGroups) -> np.ndarray:
    """sequential per-group reduction over the oracle's groups (checker of rk_group_statistics)"""
    from repkiller_b200.capi import GROUP_STATS_DTYPE
    L = lib()
    L.rko_group_statistics.argtypes = [C.c_void_p, C.POINTER(_Result), C.c_void_p]
    L.rko_group_statistics.restype = C.c_int
    r = _Result()
    r.n_kept, r.n_groups = g.n_kept, g.n_groups
    keep = [np.ascontiguousarray(g.order), np.ascontiguousarray(g.out_gid), np.ascontiguousarray(g.identity)]
    r.order = keep[0].ctypes.data_as(C.POINTER(C.c_uint32))
    r.out_gid = keep[1].ctypes.data_as(C.POINTER(C.c_uint32))
    r.identity = keep[2].ctypes.data_as(C.POINTER(C.c_float))
    rec = np.ascontiguousarray(records)
    out = np.zeros(g.n_groups, dtype=GROUP_STATS_DTYPE)
    if L.rko_group_statistics(rec.ctypes.data, C.byref(r), out.ctypes.data):
        raise ValueError("rko_group_statistics failed")
    return out


def std_sort_by_key(idx: np.ndarray, h: np.ndarray) -> np.ndarray:
    """libstdc++ std::sort order of idx under comp(a, b) = h[a] < h[b]."""
    out = np.ascontiguousarray(idx, dtype=np.uint32).copy()
    hh = np.ascontiguousarray(h, dtype=np.uint64)
    lib().rko_std_sort_by_key(out.ctypes.data, out.shape[0], hh.ctypes.data)
    return out


def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_ref(in_csv: str, out_csv: str, len_ratio: float, pos_ratio: float, nosave: bool = False, timeout: float | None = None) -> dict:
    """Run the compiled reference (oracle/_ref/repkiller_ref); returns its per-phase timing line."""
    env = dict(os.environ)
    if nosave:
        env["RK_REF_NOSAVE"] = "1"
    p = subprocess.run([REF_BIN, in_csv, out_csv, repr(len_ratio), repr(pos_ratio)], env=env, capture_output=True,
                       timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"repkiller_ref rc={p.returncode}: {p.stderr[-400:]!r}")
    return json.loads(p.stderr.decode().strip().splitlines()[-1])
