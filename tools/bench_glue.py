"""Times the torch glue operations of the partitioned path in isolation (tuning aid)."""
import torch, time
dev = torch.device("cuda:0")
n = 10_000_000
cols = [torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device=dev) for _ in range(7)]
perm64 = torch.randperm(n, device=dev)
perm32 = perm64.to(torch.int32)
def t(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:50s} {e0.elapsed_time(e1) / reps:8.3f} ms")
    return out
table = t("stack 7 cols -> [n,7]", lambda: torch.stack(cols, dim=1))
t("row gather table[perm64]", lambda: table[perm64])
t("row gather table[perm32]", lambda: table[perm32])
t("index_select(table,0,perm32)", lambda: torch.index_select(table, 0, perm32))
t("7 column gathers col[perm64]", lambda: [c[perm64] for c in cols])
t("7 column gathers + stack", lambda: torch.stack([c[perm64] for c in cols], dim=1))
t("table.t().contiguous()", lambda: table.t().contiguous())
t("7x table[:,j].contiguous()", lambda: [table[:, j].contiguous() for j in range(7)])
t("scatter out[perm64]=v", lambda: torch.empty_like(cols[0]).index_put_((perm64,), cols[1]))
k = torch.sort(cols[0] >> 8).values
t("bincount(k>>7)", lambda: torch.bincount((k >> 7).to(torch.int64), minlength=65536))
t("searchsorted 7 cuts", lambda: torch.searchsorted(k, torch.tensor([1, 2, 3, 4, 5, 6, 7], device=dev, dtype=torch.int32) * 1000000))
