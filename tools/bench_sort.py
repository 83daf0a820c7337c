"""Micro-benchmark of the K2 radix sort alone (tuning aid): python tools/bench_sort.py [n] [bits]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repkiller_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ctx = capi.Context(0)
dev = torch.device("cuda:0")
keys = torch.randint(0, 2 ** bits, (n,), dtype=torch.int64).to(torch.int32).to(dev)
ko, vo, kt, vt = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4))
work = torch.empty(ctx.sort_pairs_work_bytes(n), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
args = (keys.data_ptr(), None, ko.data_ptr(), vo.data_ptr(), kt.data_ptr(), vt.data_ptr(), n, bits, work.data_ptr())
for _ in range(3):
    ctx.sort_pairs_device(*args)
ctx.profile_enable(True)
ctx.profile_read(reset=True)
reps = 10
for _ in range(reps):
    ctx.sort_pairs_device(*args)
prof = ctx.profile_read()
passes = (bits + 7) // 8
for k, (l, ms, _u) in prof.items():
    print(f"{os.environ.get('RK_LIB_SUFFIX','')} {k}: {ms / reps * 1e3:.1f} us per sort, {ms / l * 1e3:.1f} us per launch")
sc = prof["k_radix_scatter"][1] / reps / passes
print(f"pass: {sc * 1e3:.1f} us -> {16 * n / sc / 1e6:.0f} GB/s algorithmic")
