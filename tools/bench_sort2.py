"""Sort micro-benchmark over key distributions (tuning aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repkiller_b200 import capi
n, bits = 10_000_000, 23
ctx = capi.Context(0)
dev = torch.device("cuda:0")
rnd = torch.randint(0, 2 ** bits, (n,), dtype=torch.int64, device=dev)
dists = {
    "random": rnd,
    "sorted": torch.sort(rnd).values,
    "nearly sorted (+-64)": torch.clamp(torch.sort(rnd).values + torch.randint(-64, 64, (n,), device=dev), 0, 2 ** bits - 1),
    "gid-like (0.64*i + noise)": torch.clamp((torch.arange(n, device=dev) * 0.64).long() - torch.randint(0, 5000, (n,), device=dev), 0, 2 ** bits - 1),
    "constant": torch.zeros(n, dtype=torch.int64, device=dev),
}
ko, vo, kt, vt = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4))
work = torch.empty(ctx.sort_pairs_work_bytes(n), dtype=torch.uint8, device=dev)
for name, k in dists.items():
    keys = k.to(torch.int32)
    args = (keys.data_ptr(), None, ko.data_ptr(), vo.data_ptr(), kt.data_ptr(), vt.data_ptr(), n, bits, work.data_ptr())
    for _ in range(2):
        ctx.sort_pairs_device(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.profile_enable(True); ctx.profile_read(reset=True)
    e0.record()
    for _ in range(5):
        ctx.sort_pairs_device(*args)
    e1.record(); torch.cuda.synchronize()
    prof = ctx.profile_read(); ctx.profile_enable(False)
    print(f"{name:28s} wall {e0.elapsed_time(e1)/5:7.3f} ms   " + "  ".join(f"{kk}={v[1]/5:.3f}" for kk, v in prof.items()))
