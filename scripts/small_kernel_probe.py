"""Warm-cache CUDA-event timing of the question-side kernels at the cfg2 shapes (B=64, Q=12, D=200, I=2, N=2000).
ncu's per-launch times are cold-cache (it flushes L2 before every kernel); these are the in-loop numbers."""
import json
import sys

import torch

sys.path.insert(0, ".")
from gnn_rag_b200 import ops  # noqa: E402

dev = "cuda"
B, Q, D, I, N, W = 64, 12, 200, 2, 2000, 300
torch.manual_seed(0)


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n


res = {}
lstm = torch.nn.LSTM(W, D, batch_first=True).to(dev)
x = torch.randn(B, Q, W, device=dev)
with torch.no_grad():
    gx = torch.nn.functional.linear(x, lstm.weight_ih_l0, lstm.bias_ih_l0)
    res["lstm_forward_us"] = timeit(lambda: ops.lstm_forward(gx, lstm.weight_hh_l0, lstm.bias_hh_l0))
    z = torch.zeros(1, B, D, device=dev)
    res["cudnn_lstm_us"] = timeit(lambda: lstm(x, (z, z)))
    hidden = torch.randn(B, Q, D, device=dev)
    qn = torch.randn(B, D, device=dev)
    text = torch.randint(0, 30, (B, Q), device=dev)
    Wq = [torch.randn(D, D, device=dev) * 0.05 for _ in range(I)]
    bq = [torch.randn(D, device=dev) for _ in range(I)]
    Wcq, bcq = torch.randn(D, 4 * D, device=dev) * 0.05, torch.randn(D, device=dev)
    wca, bca = torch.randn(D, device=dev), torch.randn(1, device=dev)
    res["instructions_us"] = timeit(lambda: ops.instructions(hidden, qn, text, 30, Wq, bq, Wcq, bcq, wca, bca))
    h = torch.randn(B * N, D, device=dev)
    seed = torch.zeros(B, N, device=dev)
    seed[:, 3] = 1.0
    ins = torch.randn(B, I, D, device=dev)
    Wr = [torch.randn(D, 3 * D, device=dev) * 0.05 for _ in range(I)]
    Wg = [torch.randn(D, 3 * D, device=dev) * 0.05 for _ in range(I)]
    res["query_reform_us"] = timeit(lambda: ops.query_reform(seed, h, ins, Wr, Wg, B, N))
    dist = torch.softmax(torch.randn(B, N, device=dev), 1)
    teacher = torch.zeros(B, N, device=dev)
    teacher[:, 5] = 1
    res["kl_loss_pred_us"] = timeit(lambda: ops.kl_loss_pred(dist, teacher))
    dots = torch.randn(2 * B * N, device=dev)
    mask = torch.ones(B * N, device=dev)
    bsc = torch.zeros(1, device=dev)
    res["masked_softmax_us"] = timeit(lambda: ops.masked_softmax(dots, bsc, mask, B, N))
print(json.dumps(res, indent=1))
