"""Aggregation kernel variants at the cfg2 hot shape: bit-equality against the round-1 persistent kernel and
CUDA-event timing (L2 flushed between launches).   python scripts/agg_modes_probe.py [modes...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_rag_b200 import batching, ops, synthetic as S  # noqa: E402

dev = torch.device("cuda")
c = S.CONFIGS["cfg2"]
B, N, D, I = c["B"], c["N"], c["D"], c["I"]
R = S.WEBQSP_NUM_RELATION
batch = S.make_batch(1, B=B, N=N, E=c["E"], with_weights=False)
db = batching.stage_batch(batch, dev, R + 1)
g = db.graph
F = g.F
rs = np.random.RandomState(0)
tab = torch.from_numpy(rs.randn(2 * (R + 1), D).astype(np.float32)).to(dev)
ins = torch.from_numpy(rs.randn(B, I, D).astype(np.float32)).to(dev)
prior = torch.softmax(torch.from_numpy(rs.randn(B, N).astype(np.float32)), 1).to(dev)
pn = ops.pad_table256(tab)
pf, pi = pn[: R + 1], pn[R + 1:]
Kp = (208 * (2 * I + 1) + 63) // 64 * 64
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
abytes = 2 * F * 8 + 2 * (B * N + 1) * 4 + B * N * 4 + 2 * (R + 1) * D * 4 + B * I * D * 4 + 2 * I * B * N * D * 4


def run(mode, hot, planes):
    ops.set_option("agg_abs_ws", mode)
    ops.aggregate_dual_abs(g, prior, pf, pi, ins, planes, 208, 208)


def timeit(mode, hot, planes, n=20):
    for _ in range(3):
        run(mode, hot, planes)
    ts = []
    for _ in range(n):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(mode, hot, planes); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


modes = [int(x) for x in sys.argv[1:]] or [0, 1, 2, 3]
ref = [torch.full((B * N, Kp), 7.0, dtype=torch.bfloat16, device=dev) for _ in range(2)]
run(1, -1, tuple(ref))
torch.cuda.synchronize()
res = {}
for mode in modes:
    for hot in [-1]:
        got = [torch.full((B * N, Kp), 7.0, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        run(mode, hot, tuple(got))
        torch.cuda.synchronize()
        same = bool(torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]))
        med, best = timeit(mode, hot, tuple(got))
        res["mode%d_hot%d" % (mode, hot)] = dict(bit_equal=same, us_med=med, us_min=best,
                                                GBps=abytes / med / 1e3, frac=abytes / med / 1e3 / 6568.0)
        print("mode %d hot %5d  bit-equal %s  %.1f us (min %.1f)  %.0f GB/s  frac %.3f" % (
            mode, hot, same, med, best, abytes / med / 1e3, abytes / med / 1e3 / 6568.0), flush=True)
ops.set_option("agg_abs_ws", 2)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/agg_modes_probe.json", "w"), indent=1)
