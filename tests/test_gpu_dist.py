"""The range-partitioned multi-rank path (repkiller_b200/dist.py) with the CUDA stages, against the oracle:
one process; two processes sharing cuda:0 (gloo, host-staged exchange); and, where the box has >= 2 GPUs, two
NCCL ranks with one GPU each."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from repkiller_b200 import gen

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _workload(name):
    from dataclasses import replace
    if name == "dense":
        return replace(gen.WORKLOADS["c1"], n=60_000, lx=400_000, ly=300_000, families=40, p_rep=0.6, seed=31)
    return gen.scaled(gen.WORKLOADS["c2"], 300_000)


def _run_rank(rank, world, backend, wl, out_dir):
    from repkiller_b200 import capi
    from repkiller_b200.dist import Comm, CudaStages, group_partitioned
    devidx = rank if backend == "nccl" else 0
    torch.cuda.set_device(devidx)
    dev = torch.device("cuda", devidx)
    w = _workload(wl)
    lo, hi = w.n * rank // world, w.n * (rank + 1) // world
    lo, hi = lo - lo % 16, (hi - hi % 16 if rank + 1 < world else hi)   # slices start on a 16-record boundary
    rec = gen.generate(w, start=lo, count=hi - lo)
    aos = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()).to(dev)
    ctx = capi.Context(devidx)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    res = group_partitioned(CudaStages(ctx, dev), Comm(), aos, hi - lo, lo, w.lx + 1, w.ly + 1, w.len_ratio, w.pos_ratio)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), order=res.order.cpu().numpy().view(np.uint32),
             gid=res.gid.cpu().numpy().view(np.uint32), repval=res.repval.cpu().numpy(), identity=res.identity.cpu().numpy(),
             n_groups=res.n_groups, n_kept=res.n_kept)
    ctx.close()


def _worker(rank, world, port, backend, wl, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    _run_rank(rank, world, backend, wl, out_dir)
    dist.barrier()
    dist.destroy_process_group()


def _check(out_dir, world, wl):
    w = _workload(wl)
    rec = gen.generate(w)
    g = O.group(rec, w.lx + 1, w.ly + 1, w.len_ratio, w.pos_ratio)
    parts = [np.load(os.path.join(out_dir, f"r{r}.npz")) for r in range(world)]
    assert int(parts[0]["n_groups"]) == g.n_groups and int(parts[0]["n_kept"]) == g.n_kept
    for name, want in (("order", g.order), ("gid", g.out_gid), ("repval", g.repval)):
        got = np.concatenate([p[name] for p in parts])
        d = np.nonzero(got != want)[0]
        assert d.size == 0, f"{name}: {d.size} diffs, first at {d[:3]}"
    got = np.concatenate([p["identity"] for p in parts])
    assert np.array_equal(got.view(np.uint32), g.identity.view(np.uint32))


@pytest.mark.parametrize("wl", ["dense", "c2"])
def test_stages_single_process(tmp_path, wl):
    _run_rank(0, 1, "none", wl, str(tmp_path))
    _check(str(tmp_path), 1, wl)


@pytest.mark.parametrize("wl", ["dense", "c2"])
def test_two_ranks_one_gpu_gloo(tmp_path, wl):
    mp.spawn(_worker, args=(2, _free_port(), "gloo", wl, str(tmp_path)), nprocs=2, join=True)
    _check(str(tmp_path), 2, wl)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_ranks(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_worker, args=(world, _free_port(), "nccl", "c2", str(tmp_path)), nprocs=world, join=True)
    _check(str(tmp_path), world, "c2")
