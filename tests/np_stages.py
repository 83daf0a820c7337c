"""TEST INFRASTRUCTURE.  A small CPU re-statement of the rk_st_* stages (numpy / plain Python, sizes of a few
thousand fragments) so that the multi-rank orchestration of repkiller_b200/dist.py can be exercised on gloo
without a GPU.  Semantics follow SURVEY.md Appendix A; the whole pipeline is then compared with the oracle."""
import numpy as np
import torch

from oracle import oracle as O
from repkiller_b200.frags import FRAG_DTYPE

NONE = 0xFFFFFFFF


def _u32(t):
    return t.numpy().view(np.uint32) if isinstance(t, torch.Tensor) else np.asarray(t, dtype=np.uint32)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32 else np.ascontiguousarray(a))


def _probes_prev(c):
    return (c % 100 <= 1) and c >= 100


def _probes_next(c, M):
    r = c % 100
    if r == 99:
        return c < M or M == 0
    if r == 98:
        return M == 0 or c < M - 1
    return False


def link_words(seq_len):
    return (2 * (seq_len // 100 + 2) + 31) // 32 + 1


class NumpyStages:
    device = torch.device("cpu")

    def pack(self, cols, idx):
        return torch.stack([c.to(torch.int32) for c in cols], dim=1)[idx.to(torch.int64)]

    def unpack(self, rows, idx, want):
        src = rows if idx is None else rows[idx.to(torch.int64)]
        return [src[:, j].contiguous() for j in want]

    def scatter(self, values, idx):
        out = torch.empty_like(values)
        out[idx.to(torch.int64)] = values
        return out

    def decode(self, aos, n, lx1, ly1):
        rec = aos.numpy().view(FRAG_DTYPE)[:n]
        xs, ys, ln = (rec[k].astype(np.uint64) for k in ("xStart", "yStart", "length"))
        vsize = 1 + lx1 // 10
        key0 = (xs // 10).astype(np.uint32)
        assert (key0 < vsize).all()
        dropped = key0 == vsize - 1
        flags = ((rec["strand"] != b"f").astype(np.uint8)) | (dropped.astype(np.uint8) << 1)
        mx, my = lx1 // 100, ly1 // 100
        nbx, nby = mx + 2, my + 2
        lx = np.zeros(link_words(lx1), np.uint32)
        ly = np.zeros(link_words(ly1), np.uint32)
        for i in np.nonzero(~dropped)[0]:
            sc = int(flags[i] & 1)
            for c, M, nb, bm in ((int(xs[i] + ln[i] // 2), mx, nbx, lx), (int(ys[i] + ln[i] // 2), my, nby, ly)):
                k = sc * nb + c // 100
                if _probes_prev(c):
                    bm[k >> 5] |= np.uint32(1 << (k & 31))
                if _probes_next(c, M):
                    bm[(k + 1) >> 5] |= np.uint32(1 << ((k + 1) & 31))
        with np.errstate(all="ignore"):
            ident = (rec["ident"].astype(np.float32) * np.float32(100)) / rec["length"].astype(np.float32)
        ident = np.where(np.isnan(ident), np.frombuffer(np.uint32(0xFFC00000).tobytes(), np.float32)[0], ident).astype(np.float32)
        return dict(xs=_t(xs.astype(np.uint32)), ys=_t(ys.astype(np.uint32)), len=_t(ln.astype(np.uint32)),
                    flags=torch.from_numpy(flags), identity=torch.from_numpy(ident), key0=_t(key0), link_x=_t(lx), link_y=_t(ly),
                    n_dropped=int(dropped.sum()))

    def or_words(self, dst, src):
        dst |= src

    def sort_pairs(self, keys, bits):
        k = _u32(keys)
        perm = np.argsort(k, kind="stable")
        return _t(k[perm]), torch.from_numpy(perm.astype(np.int64))

    @staticmethod
    def _run_start(bm, k):
        while (bm[k >> 5] >> np.uint32(k & 31)) & 1:
            k -= 1
        return k

    def keys(self, m, lx1, ly1, xs_r, ys_r, len_r, flags_r, link_x, link_y):
        xs, ys, ln = _u32(xs_r).astype(np.int64), _u32(ys_r).astype(np.int64), _u32(len_r).astype(np.int64)
        bx, by = _u32(link_x), _u32(link_y)
        nbx, nby = lx1 // 100 + 2, ly1 // 100 + 2
        cx, cy = xs + ln // 2, ys + ln // 2
        sc = (flags_r.numpy() & 1).astype(np.int64)
        kx = np.array([self._run_start(bx, int(sc[i] * nbx + cx[i] // 100)) for i in range(m)], dtype=np.uint32)
        ky = np.array([self._run_start(by, int(sc[i] * nby + cy[i] // 100)) for i in range(m)], dtype=np.uint32)
        return _t(cx.astype(np.uint32)), _t(cy.astype(np.uint32)), _t(kx), _t(ky)

    @staticmethod
    def _deviation(ec, el, c, ln, lr, pr):
        if ln == 0:
            return float("nan")
        sl = -abs(abs(ln - el) / (ln * lr)) + 1.0
        if sl < 0:
            return 0.0
        sp = -abs(abs(c - ec) / (ln * pr)) + 1.0
        if sp < 0:
            return 0.0
        return sl * 0.4 + sp * 0.6

    def match(self, skey, sid, sc, slen, sxm, seq_len, lr, pr):
        key, ids, c, ln = _u32(skey), _u32(sid), _u32(sc).astype(np.int64), _u32(slen).astype(np.int64)
        xm = sxm.numpy() if sxm is not None else np.zeros(len(key), np.uint8)
        M = seq_len // 100
        owner = np.full(len(key), NONE, dtype=np.uint32)
        i = 0
        while i < len(key):
            j = i
            entries = []  # inserted positions, oldest first
            while j < len(key) and key[j] == key[i]:
                if xm[j]:
                    entries.append(j)
                    j += 1
                    continue
                cj, lj = int(c[j]), int(ln[j])
                b = cj // 100
                nbk = b - 1 if _probes_prev(cj) else (b + 1 if _probes_next(cj, M) else -1)
                best, best_sc, best_own = -1, 0.0, False
                for k in reversed(entries):   # newest first, own bucket before the neighbour, strict >
                    bk = int(c[k]) // 100
                    own = bk == b
                    if not own and bk != nbk:
                        continue
                    s = self._deviation(int(c[k]), int(ln[k]), cj, lj, lr, pr)
                    if s > best_sc or (s == best_sc and best >= 0 and own and not best_own):
                        best, best_sc, best_own = k, s, own
                if best >= 0:
                    owner[j] = ids[best]
                else:
                    entries.append(j)
                j += 1
            i = j
        return _t(owner)

    def forest(self, parent_full, m_total, lo, cnt):
        par = _u32(parent_full)
        roots = par == NONE
        gid_of_root = np.cumsum(roots) - 1
        out = np.zeros(cnt, np.uint32)
        for t in range(cnt):
            r = lo + t
            while par[r] != NONE:
                r = par[r]
            out[t] = gid_of_root[r]
        return _t(out), int(roots.sum())

    def hkey(self, k0_r, ys_r):
        k0, ys = _u32(k0_r), _u32(ys_r).astype(np.int64)
        h = np.zeros(len(k0), np.uint32)
        i = 0
        while i < len(k0):
            j = i
            while j < len(k0) and k0[j] == k0[i]:
                j += 1
            h[i:j] = np.abs(ys[i:j] - ys[j - 1]).astype(np.uint32)
            i = j
        return _t(h)

    def order(self, sgid, sh, sfidx, sident, do_sort=True):
        gid, h, fidx = _u32(sgid), _u32(sh).astype(np.uint64), _u32(sfidx)
        ident = sident.numpy()
        m = len(gid)
        src = np.arange(m, dtype=np.uint32)
        rep = np.zeros(m, np.uint8)
        i = 0
        while i < m:
            j = i
            while j < m and gid[j] == gid[i]:
                j += 1
            if j - i > 1:
                if do_sort:
                    src[i:j] = O.std_sort_by_key(src[i:j], h)
                rep[i] = 1
                rep[i + 1:j] = 2
            i = j
        return _t(fidx[src]), _t(gid.copy()), torch.from_numpy(rep), torch.from_numpy(ident[src].copy())
