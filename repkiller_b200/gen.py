"""Counter-based synthetic GECKO fragment generator ([survey choice]).

Every field of fragment ``i`` is a pure function of ``(seed, i)`` through splitmix64, so any slice of a
workload can be produced independently (per rank, per chunk) and reproduced bit-for-bit.
Records are emitted the way the reference holds them after parsing the equivalent CSV row
(/root/reference/src/FragmentsDatabase.cpp:30-43): ``diag = xStart - yStart``, ``ident = trunc(similarity)``,
``seqX = 0``, ``seqY = 1``, ``evalue = 0``.

Workload shapes (BASELINE.json ``configs``):
  c1  100k fragments, 5 Mbp x 5 Mbp, 30 % repeat-family fragments
  c2  10M fragments, 150 Mbp x 150 Mbp, 30 % repeats in 5e4 families
  c3  10M fragments, 150 Mbp, 90 % repeats in 200 families, half of them tandem
  c5  1e9 fragments, 3 Gbp x 3 Gbp (generated per partition)
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np

from .frags import FRAG_DTYPE

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_FAM = np.uint64(0xD1B54A32D192ED03)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


@dataclass(frozen=True)
class Workload:
    name: str
    n: int
    lx: int  # value printed in the CSV header (the reference loads it as lx + 1)
    ly: int
    p_rep: float
    families: int
    ax: int = 6
    ay: int = 6
    tandem_every: int = 0  # 0 = none; k = every k-th family has anchors spaced one family length apart
    seed: int = 1
    len_ratio: float = 0.05
    pos_ratio: float = 0.05


WORKLOADS = {
    "c1": Workload("c1", 100_000, 5_000_000, 5_000_000, 0.3, 500, seed=1),
    "c2": Workload("c2", 10_000_000, 150_000_000, 150_000_000, 0.3, 50_000, seed=2),
    "c3": Workload("c3", 10_000_000, 150_000_000, 150_000_000, 0.9, 200, tandem_every=2, seed=3),
    # config 4 (5,000-contig draft assembly vs a 150 Mbp reference, 50M fragments).  The reference reads every row as
    # sequence pair (0, 1) — seqX/seqY are hard-coded, FragmentsDatabase.cpp:41-42 — so a multi-contig file is ONE
    # comparison over the concatenated X coordinates: contigs with log-uniform lengths 2 kbp..2 Mbp add up to about
    # 1.45 Gbp, fragments land on a contig in proportion to its length, i.e. uniformly on the concatenated axis.
    "c4": Workload("c4", 50_000_000, 1_450_000_000, 150_000_000, 0.3, 250_000, seed=4),
    "c5": Workload("c5", 1_000_000_000, 3_000_000_000, 3_000_000_000, 0.3, 5_000_000, seed=5),
}


def scaled(w: Workload, n: int) -> Workload:
    """Same shape at a different fragment count: sequence lengths and family count scale with n so the
    per-bucket density (and so the work per fragment) stays that of the named workload."""
    f = n / w.n
    return replace(w, name=f"{w.name}@{n}", n=n, lx=max(20_000, int(w.lx * f)), ly=max(20_000, int(w.ly * f)),
                   families=max(1, int(w.families * f)))


def _u(seed: int, i: np.ndarray, k: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(seed) * _GOLD + np.uint64(16) * i + np.uint64(k))


def _fam(seed: int, j: np.ndarray, t) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64((np.uint64(seed) * _GOLD) ^ (_FAM + np.uint64(64) * j + np.asarray(t, dtype=np.uint64)))


def generate(w: Workload, start: int = 0, count: int | None = None) -> np.ndarray:
    """Records ``start .. start+count`` of workload ``w`` (file order)."""
    if count is None:
        count = w.n - start
    out = np.zeros(count, dtype=FRAG_DTYPE)
    chunk = 1 << 20
    for c0 in range(0, count, chunk):
        c1 = min(count, c0 + chunk)
        _fill(w, np.arange(start + c0, start + c1, dtype=np.uint64), out[c0:c1])
    return out


def _fill(w: Workload, i: np.ndarray, out: np.ndarray) -> None:
    s = w.seed
    u64 = np.uint64
    is_rep = (_u(s, i, 0) >> u64(11)).astype(np.float64) * (1.0 / (1 << 53)) < w.p_rep

    # background fragments
    bl = u64(40) + _u(s, i, 7) % u64(2961)
    bx = _u(s, i, 8) % (u64(w.lx) - bl - u64(1))
    by = _u(s, i, 9) % (u64(w.ly) - bl - u64(1))

    # repeat-family copies: family length, one of ax (ay) anchors per axis, small jitter
    j = _u(s, i, 1) % u64(w.families)
    fl = u64(40) + _fam(s, j, 0) % u64(1961)
    a = _u(s, i, 2) % u64(w.ax)
    b = _u(s, i, 3) % u64(w.ay)
    span_x = u64(w.lx) - fl * u64(w.ax + 1) - u64(32)
    span_y = u64(w.ly) - fl * u64(w.ay + 1) - u64(32)
    if w.tandem_every:
        tandem = (j % u64(w.tandem_every)) == 0
    else:
        tandem = np.zeros(i.shape, dtype=bool)
    ax_free = u64(8) + _fam(s, j, u64(1) + a) % span_x
    ay_free = u64(8) + _fam(s, j, u64(17) + b) % span_y
    ax_tan = u64(8) + _fam(s, j, 1) % span_x + a * fl
    ay_tan = u64(8) + _fam(s, j, 17) % span_y + b * fl
    xa = np.where(tandem, ax_tan, ax_free)
    ya = np.where(tandem, ay_tan, ay_free)
    dx = _u(s, i, 4) % u64(11)
    dy = _u(s, i, 5) % u64(11)
    dl = _u(s, i, 6) % u64(7)
    rl = fl + dl - u64(3)
    rx = xa + dx - u64(5)
    ry = ya + dy - u64(5)

    length = np.where(is_rep, rl, bl)
    xs = np.where(is_rep, rx, bx)
    ys = np.where(is_rep, ry, by)
    strand = np.where((_u(s, i, 10) & u64(1)) == 0, b"f", b"r")
    frac = 0.65 + 0.35 * ((_u(s, i, 11) >> u64(11)).astype(np.float64) * (1.0 / (1 << 53)))
    ident_true = np.floor(length.astype(np.float64) * frac).astype(np.uint64)
    sim = (100.0 * ident_true.astype(np.float64) / length.astype(np.float64)).astype(np.float32)

    out["xStart"] = xs
    out["yStart"] = ys
    out["xEnd"] = xs + length - u64(1)
    out["yEnd"] = ys + length - u64(1)
    out["length"] = length
    out["diag"] = xs.astype(np.int64) - ys.astype(np.int64)
    out["score"] = u64(4) * ident_true
    out["similarity"] = sim
    out["ident"] = sim.astype(np.uint64)  # the reference's (uint64_t) stof(similarity)
    out["seqX"] = 0
    out["seqY"] = 1
    out["block"] = 0
    out["strand"] = strand
