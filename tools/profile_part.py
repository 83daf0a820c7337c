"""Where the partitioned path spends its device time (1 process, world size 1: every exchange is a local no-op, the
rest of the machinery runs as on N GPUs).  python tools/profile_part.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from repkiller_b200 import capi, gen
from repkiller_b200.dist import Comm, CudaStages, group_partitioned
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
w = gen.scaled(gen.WORKLOADS["c2"], n)
dev = torch.device("cuda:0")
ctx = capi.Context(0)
buf = torch.empty(n * 109 + 16, dtype=torch.uint8, device=dev)
ctx.generate_device(w, 0, n, buf.data_ptr())
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
st, comm = CudaStages(ctx, dev), Comm()
with torch.cuda.stream(stream):
    for _ in range(2):
        group_partitioned(st, comm, buf, n, 0, w.lx + 1, w.ly + 1, 0.05, 0.05)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        group_partitioned(st, comm, buf, n, 0, w.lx + 1, w.ly + 1, 0.05, 0.05)
    e1.record(stream)
    torch.cuda.synchronize()
    print("ms per call", e0.elapsed_time(e1) / 3)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        group_partitioned(st, comm, buf, n, 0, w.lx + 1, w.ly + 1, 0.05, 0.05)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
