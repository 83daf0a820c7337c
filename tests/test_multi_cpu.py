"""Host-side plumbing of the partitioned path (repkiller_b200/multi.py) on CPU: the file slices, the byte-string exchange
of the bootstrap over a world_size-2 gloo group, and the additivity of the output checksum.  The partitioning itself runs
inside librk_b200.so and is covered by tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from repkiller_b200 import capi, multi


def test_slices_cover_the_file_on_16_record_boundaries():
    for n in (0, 1, 15, 16, 17, 1000, 123_457, 10_000_000):
        for world in (1, 2, 3, 5, 8, 16):
            bounds = [multi.slice_bounds(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            for (lo, hi), (lo2, _) in zip(bounds, bounds[1:]):
                assert hi == lo2 and lo <= hi
            assert all(lo % 16 == 0 for lo, _ in bounds)
            if n >= 64 * world:
                sizes = [hi - lo for lo, hi in bounds]
                assert max(sizes) - min(sizes) <= 32


def test_checksum_of_ranges_adds_up_to_the_checksum_of_the_whole():
    rng = np.random.default_rng(3)
    m = 10_000
    order = rng.permutation(m).astype(np.uint32)
    gid = np.sort(rng.integers(0, 4000, m)).astype(np.uint32)
    rep = rng.integers(0, 3, m).astype(np.uint8)
    ident = rng.random(m, dtype=np.float32) * 100
    whole = multi.output_checksum(order, gid, rep, ident)
    for cuts in ([0, m], [0, 1, m], [0, 2500, 2500, 7000, m]):
        parts = sum(multi.output_checksum(order[a:b], gid[a:b], rep[a:b], ident[a:b], a) for a, b in zip(cuts, cuts[1:])) & ((1 << 64) - 1)
        assert parts == whole
    swapped = order.copy()
    swapped[[10, 11]] = swapped[[11, 10]]
    assert multi.output_checksum(swapped, gid, rep, ident) != whole   # position dependent


class _FakeContext:
    """records what bootstrap() hands to the C ABI"""
    def __init__(self, rank):
        self.rank = rank
        self.calls = []

    def dist_init(self, rank, world, unique_id, cap):
        self.calls.append(("init", rank, world, bytes(unique_id), cap))

    def dist_export(self):
        return bytes([self.rank]) * capi.DIST_BLOB_BYTES

    def dist_import(self, blobs):
        self.calls.append(("import", bytes(blobs)))


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _FakeContext(rank)
    real = capi.dist_unique_id
    capi.dist_unique_id = lambda: bytes(range(128))   # no NCCL (no GPU) on the CPU box: a stand-in id from rank 0
    try:
        multi.bootstrap(ctx, 12345)
    finally:
        capi.dist_unique_id = real
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array(ctx.calls, dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_bootstrap_moves_the_id_and_the_blobs_over_gloo(tmp_path, world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        calls = np.load(os.path.join(tmp_path, f"r{r}.npy"), allow_pickle=True)
        init, imp = calls[0], calls[1]
        assert tuple(init[:3]) == ("init", r, world) and init[3] == bytes(range(128)) and init[4] == 12345
        assert imp[0] == "import" and imp[1] == b"".join(bytes([q]) * capi.DIST_BLOB_BYTES for q in range(world))
