"""torch.profiler view of one partitioned step (tuning aid).  torchrun --nproc-per-node 2 tools/profile_part.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from repkiller_b200 import capi, gen
from repkiller_b200.dist import Comm, CudaStages, group_partitioned
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
w = gen.scaled(gen.WORKLOADS["c2"], 10_000_000 * world)
lo, hi = w.n * rank // world, w.n * (rank + 1) // world
lo, hi = lo - lo % 16, (hi - hi % 16 if rank + 1 < world else hi)
rec = gen.generate(w, start=lo, count=hi - lo)
dev = torch.from_numpy(rec.view(np.uint8).reshape(-1)).to(device)
ctx = capi.Context(local); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
st, comm = CudaStages(ctx, device), Comm()
run = lambda: group_partitioned(st, comm, dev, hi - lo, lo, w.lx + 1, w.ly + 1, 0.05, 0.05)
for _ in range(3): run()
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))
dist.barrier(); dist.destroy_process_group()
