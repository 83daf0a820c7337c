"""One comparison range-partitioned over the GPUs of a box (BASELINE config 5).

Every rank starts with a contiguous slice of the fragment file.  The greedy grouping needs three different
orderings of the fragments, so there are three redistributions (all-to-all over NCCL/NVLink), each by contiguous
ranges of a sort key so that the unit the kernels work on never straddles two GPUs:

  1. by xStart/10 ranges      -> processing order (global rank = offset of the GPU + local position); an X bucket
                                  (the unit of generate_diagonal_func) lives on one GPU;
  2. by X super-bucket ranges -> X pass (K3); owners travel back to the fragment's home GPU (reverse all-to-all);
     by Y super-bucket ranges -> Y pass, the same way (needs the X result as the "already matched" flag);
  3. parents are all-gathered (4 B per fragment and GPU) and every GPU resolves the roots of its own slice (K4);
     group ids are global because every GPU scans the same full array;
  4. by group-id ranges       -> per-group ordering and labels (K5); GPU r ends up with the output lines of the
                                  r-th range of groups, in the reference's output order.

Super-bucket keys use link maps OR-ed over all GPUs, so a run of linked buckets has the same key everywhere and
is moved as a whole.  Ties are broken by global rank only, never by GPU id: the result is bit-identical for any
number of GPUs (checked by tests/test_dist_cpu.py on gloo and tests/test_gpu_dist.py on NCCL).

The kernels are reached through the rk_st_* entry points of the C ABI (`CudaStages`); torch provides device
buffers, index plumbing between stages and torch.distributed.  The orchestration is backend-agnostic: the CPU
tests run it over gloo with a small numpy re-statement of the stages that lives in tests/.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch
import torch.distributed as dist

NONE = -1  # 0xFFFFFFFF in the int32 bit containers used for all u32 columns


def _iota_u32(n: int, start: int, device) -> torch.Tensor:
    """start, start+1, ... as u32 values in an int32 container (two's complement wrap above 2^31)"""
    s32 = ((int(start) + 2 ** 31) % 2 ** 32) - 2 ** 31
    return torch.arange(n, dtype=torch.int32, device=device) + s32


def ceil_log2(x: int) -> int:
    """bits needed for values 0 .. x-1 (at least 1)"""
    return max(1, (max(int(x), 1) - 1).bit_length())


# ---------------------------------------------------------------------------------------------------------
# communication plumbing (nccl on GPUs, gloo in the CPU tests)
# ---------------------------------------------------------------------------------------------------------
@dataclass
class Plan:
    send_counts: list
    recv_counts: list


class Comm:
    """Collectives of one process group.  With the gloo backend device tensors are staged through host memory, so
    the multi-rank path can also be exercised with several processes on one GPU (tests); NCCL moves device
    buffers directly over NVLink."""

    def __init__(self, group=None):
        self.group = group
        on = dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.size = dist.get_world_size(group) if on else 1
        self.staged = on and dist.get_backend(group) == "gloo"
        self.bytes_sent = 0

    # -- low level: three collectives, optionally staged through the host
    def _a2a(self, out, inp, out_splits=None, in_splits=None):
        if self.staged and inp.is_cuda:
            o = torch.empty(out.shape, dtype=out.dtype)
            dist.all_to_all_single(o, inp.cpu().contiguous(), out_splits, in_splits, group=self.group)
            out.copy_(o)
        else:
            dist.all_to_all_single(out, inp.contiguous(), out_splits, in_splits, group=self.group)

    def _allgather_equal(self, t: torch.Tensor) -> torch.Tensor:
        """[size, numel] of equally sized contiguous tensors"""
        n = t.numel()
        src = t.contiguous().view(-1)
        if self.staged or not t.is_cuda:
            host = src.cpu()
            parts = [torch.empty_like(host) for _ in range(self.size)]
            dist.all_gather(parts, host, group=self.group)
            return torch.stack(parts).to(t.device)
        out = torch.empty(self.size * n, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, src, group=self.group)
        return out.view(self.size, n)

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.size > 1:
            if self.staged and t.is_cuda:
                h = t.cpu()
                dist.all_reduce(h, op=dist.ReduceOp.SUM, group=self.group)
                t.copy_(h)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    # -- what the grouping uses
    def all_gather_ints(self, value: int, device) -> list:
        if self.size == 1:
            return [int(value)]
        return [int(x) for x in self._allgather_equal(torch.tensor([int(value)], dtype=torch.int64, device=device)).view(-1).tolist()]

    def all_gather_var(self, t: torch.Tensor, counts: list) -> torch.Tensor:
        """concatenation over ranks of 1-D tensors with the given lengths"""
        if self.size == 1:
            return t
        mx = max(counts)
        pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        out = self._allgather_equal(pad)
        self.bytes_sent += pad.numel() * pad.element_size() * (self.size - 1)
        return torch.cat([out[r, : counts[r]] for r in range(self.size)])

    def all_gather_same(self, t: torch.Tensor) -> torch.Tensor:
        """[size, len(t)] of equally sized 1-D tensors"""
        return t.unsqueeze(0) if self.size == 1 else self._allgather_equal(t)

    def exchange(self, rows: torch.Tensor, send_counts: list):
        """rows: [n, k] sorted by destination; returns ([n_recv, k] in source-rank order, Plan)"""
        if self.size == 1:
            return rows, Plan(list(send_counts), list(send_counts))
        sc = torch.tensor(send_counts, dtype=torch.int64, device=rows.device)
        rc = torch.empty_like(sc)
        self._a2a(rc, sc)
        recv_counts = [int(x) for x in rc.tolist()]
        out = torch.empty((sum(recv_counts),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        self._a2a(out, rows, recv_counts, list(send_counts))
        self.bytes_sent += (rows.shape[0] - send_counts[self.rank]) * rows.element_size() * (rows.shape[1] if rows.dim() > 1 else 1)
        return out, Plan(list(send_counts), recv_counts)

    def exchange_back(self, values: torch.Tensor, plan: Plan) -> torch.Tensor:
        """values aligned with the rows an exchange() delivered; returns them aligned with the rows it sent"""
        if self.size == 1:
            return values
        out = torch.empty((sum(plan.send_counts),) + tuple(values.shape[1:]), dtype=values.dtype, device=values.device)
        self._a2a(out, values, plan.send_counts, plan.recv_counts)
        self.bytes_sent += (values.shape[0] - plan.recv_counts[self.rank]) * values.element_size()
        return out

    def range_partition(self, sorted_keys: torch.Tensor, bits: int) -> list:
        """Send counts that split the global key range into `size` contiguous ranges of about equal population.
        Cuts are key values, so equal keys always land on the same rank."""
        if self.size == 1:
            return [int(sorted_keys.shape[0])]
        shift = max(0, bits - 16)
        nb = 1 << min(bits, 16)
        # local cumulative histogram over nb coarse bins straight from the sorted keys
        edges = (torch.arange(1, nb + 1, dtype=torch.int64, device=sorted_keys.device) << shift)
        edges = torch.clamp(edges, max=2 ** 31 - 1).to(sorted_keys.dtype)
        cum = torch.searchsorted(sorted_keys, edges, right=False).to(torch.int64)
        cum[-1] = sorted_keys.shape[0]
        cum = self.all_reduce_sum(cum.contiguous())
        total = int(cum[-1].item())
        targets = torch.tensor([(total * r) // self.size for r in range(1, self.size)], dtype=torch.int64, device=cum.device)
        cut_bins = torch.searchsorted(cum, targets, right=False) + 1   # first bin boundary with cum >= target
        cut_bins = torch.clamp(cut_bins, max=nb)
        cut_keys = torch.clamp(cut_bins << shift, max=2 ** 31 - 1).to(sorted_keys.dtype)
        bounds = torch.searchsorted(sorted_keys, cut_keys, right=False).tolist() if sorted_keys.numel() else [0] * (self.size - 1)
        bounds = [0] + [int(b) for b in bounds] + [int(sorted_keys.shape[0])]
        for i in range(1, len(bounds)):
            bounds[i] = max(bounds[i], bounds[i - 1])
        return [bounds[i + 1] - bounds[i] for i in range(self.size)]


# ---------------------------------------------------------------------------------------------------------
# the CUDA stages (C ABI rk_st_*)
# ---------------------------------------------------------------------------------------------------------
def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else (C.c_void_p(t.data_ptr()) if t is not None else None)


class CudaStages:
    """Kernels K1..K5 on torch device tensors through the C ABI.  No CPU path: needs the library and a GPU."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.L = ctx._L
        self.h = ctx._h
        self.device = device
        L = self.L
        V, U64, D = C.c_void_p, C.c_uint64, C.c_double
        L.rk_st_link_words.argtypes = [U64]
        L.rk_st_link_words.restype = U64
        L.rk_st_decode.argtypes = [V, V, U64, U64, U64] + [V] * 8 + [C.POINTER(U64)]
        L.rk_st_or_words.argtypes = [V, V, V, U64]
        L.rk_st_keys.argtypes = [V, U64, U64, U64] + [V] * 10
        L.rk_st_match.argtypes = [V, U64, V, V, V, V, V, U64, D, D, V]
        L.rk_st_forest.argtypes = [V, V, U64, U64, U64, V, C.POINTER(U64)]
        L.rk_st_hkey.argtypes = [V, V, V, U64, V]
        L.rk_st_order.argtypes = [V, U64, V, V, V, V, C.c_int, V, V, V, V]
        L.rk_st_interleave.argtypes = [V, C.POINTER(V), U64, C.c_int, V]
        L.rk_st_gather_rows.argtypes = [V, V, V, U64, C.c_int, V]
        L.rk_st_unpack_rows.argtypes = [V, V, V, U64, C.c_int, C.POINTER(V)]
        L.rk_st_scatter.argtypes = [V, V, V, U64, V]

    def _i32(self, n):
        return torch.empty(max(int(n), 0), dtype=torch.int32, device=self.device)

    # -- row plumbing: a fragment travels between ranks as a row of k 32-bit words
    def pack(self, cols: list, idx: torch.Tensor) -> torch.Tensor:
        """rows[i][j] = cols[j][idx[i]] as an [len(idx), k] int32 tensor (interleave once, then one row gather)"""
        k, n = len(cols), cols[0].shape[0]
        cols = [c.contiguous() for c in cols]
        ptrs = (C.c_void_p * k)(*[c.data_ptr() for c in cols])
        table = torch.empty((n, k), dtype=torch.int32, device=self.device)
        self.ctx._check(self.L.rk_st_interleave(self.h, ptrs, n, k, _p(table)))
        out = torch.empty((idx.shape[0], k), dtype=torch.int32, device=self.device)
        self.ctx._check(self.L.rk_st_gather_rows(self.h, _p(table), _p(idx), idx.shape[0], k, _p(out)))
        return out

    def unpack(self, rows: torch.Tensor, idx, want: list) -> list:
        """columns `want` of rows[idx] (idx None: rows in place) as contiguous 1-D tensors"""
        k = rows.shape[1]
        n = idx.shape[0] if idx is not None else rows.shape[0]
        outs = {j: self._i32(n) for j in want}
        ptrs = (C.c_void_p * k)(*[outs[j].data_ptr() if j in outs else None for j in range(k)])
        self.ctx._check(self.L.rk_st_unpack_rows(self.h, _p(rows), _p(idx) if idx is not None else None, n, k, ptrs))
        return [outs[j] for j in want]

    def scatter(self, values: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """out[idx[i]] = values[i] (idx is a permutation)"""
        out = torch.empty_like(values)
        self.ctx._check(self.L.rk_st_scatter(self.h, _p(values), _p(idx), values.shape[0], _p(out)))
        return out

    def decode(self, aos: torch.Tensor, n: int, lx1: int, ly1: int) -> dict:
        xs, ys, ln, k0 = (self._i32(n) for _ in range(4))
        flags = torch.empty(n, dtype=torch.uint8, device=self.device)
        ident = torch.empty(n, dtype=torch.float32, device=self.device)
        link_x = self._i32(self.L.rk_st_link_words(lx1))
        link_y = self._i32(self.L.rk_st_link_words(ly1))
        nd = C.c_uint64(0)
        self.ctx._check(self.L.rk_st_decode(self.h, _p(aos), n, lx1, ly1, _p(xs), _p(ys), _p(ln), _p(flags), _p(ident), _p(k0),
                                            _p(link_x), _p(link_y), C.byref(nd)))
        return dict(xs=xs, ys=ys, len=ln, flags=flags, identity=ident, key0=k0, link_x=link_x, link_y=link_y, n_dropped=nd.value)

    def or_words(self, dst: torch.Tensor, src: torch.Tensor):
        self.ctx._check(self.L.rk_st_or_words(self.h, _p(dst), _p(src), dst.numel()))

    def sort_pairs(self, keys: torch.Tensor, bits: int):
        n = keys.shape[0]
        ko, vo, kt, vt = (self._i32(n) for _ in range(4))
        if n:
            work = torch.empty(self.ctx.sort_pairs_work_bytes(n), dtype=torch.uint8, device=self.device)
            self.ctx.sort_pairs_device(keys.data_ptr(), None, ko.data_ptr(), vo.data_ptr(), kt.data_ptr(), vt.data_ptr(), n,
                                       bits, work.data_ptr())
        return ko, vo   # int32 permutation (torch indexes with it directly)

    def keys(self, m, lx1, ly1, xs_r, ys_r, len_r, flags_r, link_x, link_y):
        cx, cy, kx, ky = (self._i32(m) for _ in range(4))
        self.ctx._check(self.L.rk_st_keys(self.h, m, lx1, ly1, _p(xs_r), _p(ys_r), _p(len_r), _p(flags_r), _p(link_x), _p(link_y),
                                          _p(cx), _p(cy), _p(kx), _p(ky)))
        return cx, cy, kx, ky

    def match(self, skey, sid, sc, slen, sxm, seq_len, len_ratio, pos_ratio):
        m = skey.shape[0]
        owner = self._i32(m)
        self.ctx._check(self.L.rk_st_match(self.h, m, _p(skey), _p(sid), _p(sc), _p(slen), _p(sxm) if sxm is not None else None,
                                           seq_len, len_ratio, pos_ratio, _p(owner)))
        return owner

    def forest(self, parent_full, m_total, lo, cnt):
        gid = self._i32(cnt)
        ng = C.c_uint64(0)
        self.ctx._check(self.L.rk_st_forest(self.h, _p(parent_full), m_total, lo, cnt, _p(gid), C.byref(ng)))
        return gid, ng.value

    def hkey(self, k0_r, ys_r):
        h = self._i32(k0_r.shape[0])
        self.ctx._check(self.L.rk_st_hkey(self.h, _p(k0_r), _p(ys_r), k0_r.shape[0], _p(h)))
        return h

    def order(self, sgid, sh, sfidx, sident, do_sort=True):
        m = sgid.shape[0]
        o, g = self._i32(m), self._i32(m)
        rep = torch.empty(m, dtype=torch.uint8, device=self.device)
        idn = torch.empty(m, dtype=torch.float32, device=self.device)
        self.ctx._check(self.L.rk_st_order(self.h, m, _p(sgid), _p(sh), _p(sfidx), _p(sident), int(do_sort), _p(o), _p(g), _p(rep),
                                           _p(idn)))
        return o, g, rep, idn


# ---------------------------------------------------------------------------------------------------------
# the partitioned grouping
# ---------------------------------------------------------------------------------------------------------
@dataclass
class PartResult:
    """This rank's part of the output (the r-th contiguous range of groups), in the reference's output order."""
    order: torch.Tensor      # GLOBAL file index of each output line
    gid: torch.Tensor
    repval: torch.Tensor
    identity: torch.Tensor
    n_groups: int            # global
    n_kept: int              # global
    n_local_lines: int
    bytes_exchanged: int = 0
    info: dict = field(default_factory=dict)


def group_partitioned(st, comm: Comm, aos: torch.Tensor, n_local: int, file_offset: int, lx1: int, ly1: int,
                      len_ratio: float, pos_ratio: float, do_sort: bool = True, timings: dict | None = None) -> PartResult:
    """aos: this rank's records (uint8 tensor, n_local * 109 bytes, file order); file_offset: index of its first record.
    timings: if given (CUDA only), appended with the milliseconds of each section (CUDA events on the current stream)."""
    dev = aos.device
    marks = []

    def mark(name):
        if timings is not None and dev.type == "cuda":
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    vsize = 1 + lx1 // 10
    bits0 = ceil_log2(vsize)
    bitsx = ceil_log2(2 * (lx1 // 100 + 2))
    bitsy = ceil_log2(2 * (ly1 // 100 + 2))
    i32 = torch.int32

    # -- K1 on the local slice; link maps OR-ed over the ranks
    d = st.decode(aos, n_local, lx1, ly1)
    link_x, link_y = d["link_x"], d["link_y"]
    if comm.size > 1:
        for name in ("link_x", "link_y"):
            allm = comm.all_gather_same(d[name])
            acc = allm[0].clone()
            for r in range(1, comm.size):
                st.or_words(acc, allm[r].contiguous())
            if name == "link_x":
                link_x = acc
            else:
                link_y = acc
    kept = n_local - d["n_dropped"]

    mark("decode+links")
    # -- redistribution 1: processing order (stable by xStart/10, file order inside a bucket)
    k0s, lidx = st.sort_pairs(d["key0"], bits0)
    k0s, lidx = k0s[:kept], lidx[:kept]          # the dropped last X bucket carries the largest key: it sorts last
    # one 28-byte row per fragment: build the rows once (coalesced), then ONE row gather into send order
    gf = _iota_u32(n_local, file_offset, dev)
    rows = st.pack([d["key0"], gf, d["xs"], d["ys"], d["len"], d["flags"].to(i32), d["identity"].view(i32)], lidx)
    recv, _ = comm.exchange(rows, comm.range_partition(k0s, bits0))
    (rkeys,) = st.unpack(recv, None, [0])
    rk0, perm = st.sort_pairs(rkeys, bits0)                     # sources arrive in file order: stable sort = global order
    k0_r, gfidx, xs_r, ys_r, len_r, fl32, ident_r = st.unpack(recv, perm, [0, 1, 2, 3, 4, 5, 6])
    flags_r = fl32.to(torch.uint8)
    m = k0_r.shape[0]
    counts = comm.all_gather_ints(m, dev)
    off, m_total = sum(counts[: comm.rank]), sum(counts)
    grank = _iota_u32(m, off, dev)

    mark("redistribute:rank")
    # -- K2 keys
    cx, cy, kx, ky = st.keys(m, lx1, ly1, xs_r, ys_r, len_r, flags_r, link_x, link_y)

    # -- redistributions 2a/2b: one axis pass each, owners sent back home
    def axis_pass(key, c, xm, seq_len, bits, tag):
        ks, p = st.sort_pairs(key, bits)
        rows = st.pack([key, grank, c, len_r] + ([xm] if xm is not None else []), p)
        mark(tag + ":sort+pack")
        send_counts = comm.range_partition(ks, bits)
        mark(tag + ":partition")
        rcv, plan = comm.exchange(rows, send_counts)
        mark(tag + ":all_to_all")
        (rkeys,) = st.unpack(rcv, None, [0])
        rks, q = st.sort_pairs(rkeys, bits)                     # sources arrive in rank order: stable sort keeps it
        if xm is not None:
            sid, sc, slen, sx = st.unpack(rcv, q, [1, 2, 3, 4])
            sxm = sx.to(torch.uint8)
        else:
            sid, sc, slen = st.unpack(rcv, q, [1, 2, 3])
            sxm = None
        mark(tag + ":sort+unpack")
        owner_sorted = st.match(rks, sid, sc, slen, sxm, seq_len, len_ratio, pos_ratio)
        mark(tag + ":match")
        back = comm.exchange_back(st.scatter(owner_sorted, q), plan)
        owner = st.scatter(back, p)
        mark(tag + ":owners_back")
        return owner

    mark("keys")
    xo = axis_pass(kx, cx, None, lx1, bitsx, "x")
    xmatched = xo != NONE
    yo = axis_pass(ky, cy, xmatched.to(i32), ly1, bitsy, "y")
    parent = torch.where(xmatched, xo, yo)

    # -- K4 on the all-gathered forest
    parent_full = comm.all_gather_var(parent, counts)
    gid, n_groups = st.forest(parent_full.contiguous(), m_total, off, m)

    mark("forest")
    # -- K5a at home (an X bucket never straddles ranks)
    h = st.hkey(k0_r, ys_r)

    # -- redistribution 3: by group id; K5b/c
    bitsg = ceil_log2(max(n_groups, 1))
    gs, p = st.sort_pairs(gid, bitsg)
    rows = st.pack([gid, h, gfidx, ident_r], p)   # members arrive in rank order: the rank itself need not travel
    recv, _ = comm.exchange(rows, comm.range_partition(gs, bitsg))
    (rkeys,) = st.unpack(recv, None, [0])
    rgs, q = st.sort_pairs(rkeys, bitsg)
    sh, sf, si = st.unpack(recv, q, [1, 2, 3])
    o, g, rep, idn = st.order(rgs, sh, sf, si.view(torch.float32), do_sort)
    mark("redistribute:gid+order")
    if marks:
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
            timings.setdefault(n1, []).append(e0.elapsed_time(e1))
    return PartResult(o, g, rep, idn, int(n_groups), int(m_total), int(o.shape[0]), comm.bytes_sent,
                      {"m_local": m, "rank_offset": off})
