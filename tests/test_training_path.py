"""``model(batch, training=True)`` -- the call of Trainer_KBQA.train_epoch (gnn/train_model.py:222) -- against loss,
train-time metrics and parameter gradients recorded from the UNMODIFIED reference (tests/golden/train/*.npz, made by
tests/golden/make_train_golden.py).  The differentiable path is torch autograd around the aggregation kernels; the CPU
tests here switch ``autograd_path.HOST_CHECK`` on to evaluate its torch restatement on CPU tensors (the product refuses
a CPU model), the ``gpu`` tests below run it through the kernels."""
import os

import numpy as np
import pytest
import torch

import gnn_rag_b200 as G
from golden_io import GOLDEN_DIR, Golden

from gnn_rag_b200 import autograd_path


@pytest.fixture(autouse=True)
def _host_check():
    old = autograd_path.HOST_CHECK
    autograd_path.HOST_CHECK = True
    yield
    autograd_path.HOST_CHECK = old


def test_training_refuses_a_cpu_model_outside_the_host_check():
    autograd_path.HOST_CHECK = False
    m, batch, _ = _load("rearev_small")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(batch, training=True)


CASES = ["rearev_small", "rearev_norm", "rearev_posemb", "rearev_sharp_ties", "nsm_small", "nsm_reason_kb",
         "rearev_sbert_reltext"]


def _load(name, device="cpu"):
    g = Golden(name)
    t = np.load(os.path.join(GOLDEN_DIR, "train", name + ".npz"))
    args = dict(g.args, use_cuda=(device != "cpu"))
    cls = G.ReaRev if args["model_name"] == "ReaRev" else G.NSM
    m = cls(args, g.num_entity, g.num_relation, g.num_word)
    m.load_state_dict(g.sd, strict=True)
    m.eval()                                            # dropout = identity, as in the generator
    if g.rel_texts is not None:
        m.encode_rel_texts(g.rel_texts, g.rel_texts_inv)
    batch = list(g.batch[:7])
    batch[6] = t["answer_dist"]
    return m, tuple(batch), t


@pytest.mark.parametrize("name", CASES)
def test_training_forward_backward_matches_reference(name):
    m, batch, t = _load(name)
    loss, pred, pred_dist, tp_list = m(batch, training=True)
    assert abs(float(loss.detach()) - float(t["loss"])) <= 2e-5 * abs(float(t["loss"]))
    np.testing.assert_allclose(pred_dist.detach().numpy(), t["pred_dist"], rtol=2e-4, atol=1e-9)
    assert torch.equal(pred, pred_dist.argmax(1))
    h1, f1 = tp_list                                   # rearev.py:238-241: [h1.tolist(), f1.tolist()]
    assert h1 == t["h1"].tolist()
    np.testing.assert_allclose(np.array(f1), t["f1"], rtol=1e-6)
    loss.backward()
    checked = 0
    for k, p in m.named_parameters():
        key = "grad/" + k
        if key not in t.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        want = t[key]
        got = p.grad.numpy() if p.grad is not None else np.zeros_like(want)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 2e-4 * scale + 5e-9, (k, np.abs(got - want).max(), scale)   # 5e-9: the
        # score bias has a mathematically zero gradient (softmax is shift invariant), both sides hold rounding noise
        checked += 1
    assert checked >= 20


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_training_on_gpu_through_the_aggregation_kernels(name):
    """Same goldens on the GPU: the aggregation forward / backward go through gr_aggregate / gr_aggregate_backward
    (csrc/aggregate_bwd.cu, autograd_path._AggregateFn); loss and every parameter gradient against the reference."""
    from gnn_rag_b200 import autograd_path
    assert autograd_path.USE_KERNELS
    m, batch, t = _load(name, device="cuda")
    m = m.cuda()
    m.train()                                           # cuDNN's LSTM backward needs training mode ...
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()                                  # ... the dropouts stay the identity, as in the generator
    if hasattr(m.instruction, "node_encoder") and not isinstance(m.instruction.node_encoder, torch.nn.LSTM):
        m.instruction.node_encoder.eval()               # HuggingFace encoder: its internal dropouts off
    if getattr(m, "rel_texts", None) is not None:
        g = Golden(name)
        m.encode_rel_texts(g.rel_texts, g.rel_texts_inv)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):     # keep cuDNN's LSTM in fp32
        loss, pred, pred_dist, tp_list = m(batch, training=True)
    assert abs(float(loss.detach()) - float(t["loss"])) <= 2e-5 * abs(float(t["loss"]))
    np.testing.assert_allclose(pred_dist.detach().cpu().numpy(), t["pred_dist"], rtol=1e-3, atol=1e-9)
    assert tp_list[0] == t["h1"].tolist()
    np.testing.assert_allclose(np.array(tp_list[1]), t["f1"], rtol=1e-6)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        loss.backward()
    # tolerance: 5e-3 of the tensor's own scale (GPU summation orders, fp32 atomics, cuDNN's LSTM) plus 1e-5 of the
    # largest gradient in the model (tensors whose true gradient is ~0, e.g. the score bias, hold only rounding noise;
    # 1e-7 absolute for the same reason when the whole model's gradients are small)
    gmax = max(float(np.abs(t[k]).max()) for k in t.files if k.startswith("grad/"))
    checked = 0
    for k, p in m.named_parameters():
        key = "grad/" + k
        if key not in t.files:
            continue
        want = t[key]
        got = p.grad.cpu().numpy() if p.grad is not None else np.zeros_like(want)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 5e-3 * scale + 1e-5 * gmax + 1e-7, (k, np.abs(got - want).max(), scale, gmax)
        checked += 1
    assert checked >= 20


@pytest.mark.gpu
def test_kernel_backward_equals_torch_backward_at_the_hot_shape():
    """D = 200, 2 instructions, hubs and ragged questions: gradients through the kernels == gradients through the
    per-fact torch ops (index_add) on the same device, and the kernel path is not slower."""
    import time
    from gnn_rag_b200 import autograd_path, synthetic as S
    args = S.model_args("ReaRev", entity_dim=200, num_iter=2, num_ins=2, num_gnn=2, use_cuda=True, linear_dropout=0.0,
                        lm_dropout=0.0)
    torch.manual_seed(0)
    m = G.ReaRev(dict(args), 3000, 40, 100).cuda()
    b = S.make_batch(61, B=6, N=300, E=1500, num_entity=3000, num_relation=40, num_word=100, powerlaw=True,
                     n_real="ragged")
    grads, times = {}, {}
    for mode in (True, False):
        autograd_path.USE_KERNELS = mode
        try:
            for rep in range(2):
                m.zero_grad()
                torch.cuda.synchronize(); t0 = time.perf_counter()
                loss = m(b, training=True)[0]
                loss.backward()
                torch.cuda.synchronize(); times[mode] = time.perf_counter() - t0
        finally:
            autograd_path.USE_KERNELS = True
        grads[mode] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    print("train step (fwd + bwd): kernels %.1f ms, torch index_add %.1f ms" % (1e3 * times[True], 1e3 * times[False]))
    assert set(grads[True]) == set(grads[False])
    for k in grads[True]:
        a, r = grads[True][k], grads[False][k]
        assert (a - r).abs().max().item() <= 2e-4 * r.abs().max().item() + 1e-9, k


def test_two_adam_steps_reduce_the_loss():
    """The INTEGRATION.md import swap leaves Trainer_KBQA.train_epoch's inner loop working
    (gnn/train_model.py:219-231): forward(training=True) -> backward -> clip -> Adam step."""
    m, batch, _ = _load("rearev_small")
    m.train()
    torch.manual_seed(0)
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=5e-3)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss, _, _, tp_list = m(batch, training=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in m.parameters()], 1.0)
        opt.step()
        losses.append(float(loss))
        assert len(tp_list) == 2 and len(tp_list[0]) == batch[0].shape[0]
    m.eval()
    final = float(m(batch, training=True)[0])
    assert final < losses[0]
