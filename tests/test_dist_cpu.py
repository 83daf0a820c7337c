"""The multi-rank orchestration of repkiller_b200/dist.py on CPU: world_size 1, 2 and 3 over gloo with the numpy
stage re-statement (tests/np_stages.py); every partitioning must reproduce the oracle's single-process result."""
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
from repkiller_b200.dist import Comm, group_partitioned


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _workload():
    from dataclasses import replace
    # dense enough for cross-bucket links, repeat groups > 16 members and ties in h
    return replace(gen.WORKLOADS["c1"], n=3000, lx=60_000, ly=50_000, families=12, p_rep=0.6, seed=21)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from np_stages import NumpyStages
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = _workload()
    lo, hi = w.n * rank // world, w.n * (rank + 1) // world
    rec = gen.generate(w, start=lo, count=hi - lo)
    aos = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy())
    res = group_partitioned(NumpyStages(), Comm(), aos, hi - lo, lo, w.lx + 1, w.ly + 1, w.len_ratio, w.pos_ratio)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), order=res.order.numpy().view(np.uint32), gid=res.gid.numpy().view(np.uint32),
             repval=res.repval.numpy(), identity=res.identity.numpy(), n_groups=res.n_groups, n_kept=res.n_kept)
    dist.barrier()
    dist.destroy_process_group()


def _check(out_dir, world):
    w = _workload()
    rec = gen.generate(w)
    g = O.group(rec, w.lx + 1, w.ly + 1, w.len_ratio, w.pos_ratio)
    parts = [np.load(os.path.join(out_dir, f"r{r}.npz")) for r in range(world)]
    assert int(parts[0]["n_groups"]) == g.n_groups and int(parts[0]["n_kept"]) == g.n_kept
    assert np.bincount(g.gid).max() > 16, "workload must exercise the introsort path"
    for name, want in (("order", g.order), ("gid", g.out_gid), ("repval", g.repval)):
        got = np.concatenate([p[name] for p in parts])
        assert np.array_equal(got, want), name
    got = np.concatenate([p["identity"] for p in parts])
    assert np.array_equal(got.view(np.uint32), g.identity.view(np.uint32))


@pytest.mark.parametrize("world", [1, 2, 3])
def test_partitioned_grouping_matches_oracle(tmp_path, world):
    if world == 1:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from np_stages import NumpyStages
        w = _workload()
        rec = gen.generate(w)
        aos = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy())
        res = group_partitioned(NumpyStages(), Comm(), aos, w.n, 0, w.lx + 1, w.ly + 1, w.len_ratio, w.pos_ratio)
        np.savez(os.path.join(tmp_path, "r0.npz"), order=res.order.numpy().view(np.uint32), gid=res.gid.numpy().view(np.uint32),
                 repval=res.repval.numpy(), identity=res.identity.numpy(), n_groups=res.n_groups, n_kept=res.n_kept)
    else:
        mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    _check(str(tmp_path), world)
