"""Which property of the aggregation kernel's write pattern limits it?  Replays row-segment zero stores with
different row pitch / start column / lane width (no edge work) and compares with flat fills."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_rag_b200 import _lib
dev = torch.device("cuda")
def timeit(f, n=15):
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
Nt = 128000
L = _lib.load(); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
buf_hi = torch.empty(Nt * 2048, dtype=torch.bfloat16, device=dev); buf_lo = torch.empty_like(buf_hi)
P = lambda t: ctypes.c_void_p(t.data_ptr())
print("%-44s %8s %10s" % ("pattern", "us", "GB/s"))
for ld, c0, nc, mode, name in [
        (1024, 200, 800, 0, "ld1024 start200 800c 8B-lanes (current)"),
        (1024, 200, 800, 1, "ld1024 start200 800c 16B-lanes"),
        (1024, 0, 800, 0, "ld1024 start0   800c 8B-lanes"),
        (1024, 0, 800, 1, "ld1024 start0   800c 16B-lanes"),
        (1024, 224, 800, 0, "ld1024 start224 800c 8B-lanes"),
        (1024, 256, 768, 1, "ld1024 start256 768c 16B-lanes"),
        (800, 0, 800, 0, "ld800  dense rows 8B-lanes"),
        (800, 0, 800, 1, "ld800  dense rows 16B-lanes"),
        (1024, 0, 1024, 1, "ld1024 full rows 16B-lanes"),
        (2048, 400, 800, 2, "1 plane ld2048 start400 1600c 16B (fp32-like)"),
        (1600, 0, 800, 2, "1 plane ld1600 dense 16B"),
]:
    nbytes = Nt * nc * 2 * (1 if mode == 2 else 2) * (2 if mode == 2 else 1)
    us = timeit(lambda: L.gr_debug_store_probe(P(buf_hi), P(buf_lo), Nt, ld, c0, nc, mode, st))
    print("%-44s %8.1f %10.0f" % (name, us, nbytes / us / 1e3))
x = torch.empty(Nt * 800, dtype=torch.float32, device=dev)
us = timeit(lambda: x.zero_()); print("%-44s %8.1f %10.0f" % ("torch fp32 zero_ 410MB", us, x.numel() * 4 / us / 1e3))
y = torch.empty(Nt * 1600, dtype=torch.bfloat16, device=dev)
us = timeit(lambda: y.zero_()); print("%-44s %8.1f %10.0f" % ("torch bf16 zero_ 410MB", us, y.numel() * 2 / us / 1e3))
