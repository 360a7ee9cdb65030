"""Micro-probe: write-only / read-only / copy HBM bandwidth on this GPU (torch kernels, CUDA events).
Context for the aggregation kernel's roofline: that kernel is ~97% output writes."""
import torch, json
dev = "cuda"
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts)//2], ts[0]
out = {}
for mb in (410, 1024, 4096):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, dtype=torch.float32, device=dev)
    b = torch.empty(n, dtype=torch.float32, device=dev)
    med, best = timeit(lambda: a.zero_())
    out["write_zero_%dMB" % mb] = {"GBps_med": mb * 1.048576 / med, "GBps_best": mb * 1.048576 / best, "ms": med}
    med, best = timeit(lambda: a.fill_(1.5))
    out["write_fill_%dMB" % mb] = {"GBps_med": mb * 1.048576 / med, "GBps_best": mb * 1.048576 / best}
    med, best = timeit(lambda: b.copy_(a))
    out["copy_%dMB" % mb] = {"GBps_med_rw": 2 * mb * 1.048576 / med, "GBps_best_rw": 2 * mb * 1.048576 / best}
    med, best = timeit(lambda: a.sum())
    out["read_sum_%dMB" % mb] = {"GBps_med": mb * 1.048576 / med, "GBps_best": mb * 1.048576 / best}
    del a, b
print(json.dumps(out, indent=1))
