"""Tensor-level wrappers over the C ABI (include/gnnrag_b200.h): torch CUDA tensors in, torch CUDA tensors
out.  torch is used for device memory and streams only; every op below launches our own sm_100a kernels.
All ops raise if handed a CPU tensor -- there is no CPU fallback on the product path."""
import ctypes
import weakref

import torch

from . import _lib

LINEAR_RELU = 1
LINEAR_EXACT_FP32 = 2
LINEAR_W_PRESPLIT = 4
LINEAR_BF16_SINGLE = 8

# bf16 ACTIVATION STORAGE (BASELINE configs[2], "bf16"): the layer-input matrix keeps its hi plane only -- the aggregation
# kernel skips the lo plane (half the output bytes) and the e2e GEMM runs ONE bf16 product instead of three.  Tables,
# accumulation, scores and softmax stay fp32.  Off by default (cfg2's contract is fp32); bench.py --config cfg3 turns it on.
ACT_BF16 = False


class _Stats:
    """Launch accounting (bench.py's ``gpu_launches``) and optional CUDA-event timing of the aggregation
    launches (bench.py's live roofline measurement).  Events are recorded on the launching stream."""

    def __init__(self):
        self.launches = 0
        self.time_agg = False
        self.agg_events = []      # (start_event, end_event, tag)
        self.time_ops = False     # bench.py's per-kernel-class share table: events around every wrapper below
        self.op_events = []       # (start_event, end_event, class tag, info)

    def reset(self):
        self.launches = 0
        self.agg_events = []
        self.op_events = []


STATS = _Stats()


class _AggTimer:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if STATS.time_agg or STATS.time_ops:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *a):
        if STATS.time_agg or STATS.time_ops:
            self.e.record()
            if STATS.time_agg:
                STATS.agg_events.append((self.s, self.e, self.tag))
            if STATS.time_ops:
                STATS.op_events.append((self.s, self.e, "aggregation", self.tag))


class _OpTimer:
    """CUDA events (on the launching stream) around one wrapper call when ``STATS.time_ops`` is on."""

    def __init__(self, cls, info=None):
        self.cls, self.info = cls, info

    def __enter__(self):
        if STATS.time_ops:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *a):
        if STATS.time_ops:
            self.e.record()
            STATS.op_events.append((self.s, self.e, self.cls, self.info))


def _L():
    return _lib.load()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _cuda(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (no CPU fallback on this path)" % name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


def pad4(n):
    return (n + 3) & ~3


def set_option(name, value):
    _lib.check(_L().gr_set_option(name.encode(), int(value)))


class CsrGraph:
    """Both destination-CSRs of one batched subgraph (device resident).

    ``*_t``: in-edges grouped by tail (forward messages, src = head); ``*_h``: grouped by head (inverse
    messages, src = tail).  ``fact_*`` maps a CSR slot back to the original fact id, so per-fact arrays
    (weights) can be permuted with :func:`gather_f32`.
    """

    def __init__(self, B, N, F, R1, device):
        self.B, self.N, self.F, self.R1 = B, N, F, R1
        Nt = B * N
        i32 = dict(dtype=torch.int32, device=device)
        self.rowptr_t = torch.empty(pad4(Nt + 1), **i32)
        self.rowptr_h = torch.empty(pad4(Nt + 1), **i32)
        Fp = max(pad4(F), 4)
        self.src_t = torch.empty(Fp, **i32)
        self.rel_t = torch.empty(Fp, **i32)
        self.fact_t = torch.empty(Fp, **i32)
        self.src_h = torch.empty(Fp, **i32)
        self.rel_h = torch.empty(Fp, **i32)
        self.fact_h = torch.empty(Fp, **i32)
        self.status = torch.zeros(1, **i32)
        self.w_t = self.w_h = None        # normalized_gnn edge weights (1/outdeg(head))
        self.wr_t = self.wr_h = None      # norm_rel edge weights (1/count(head, rel))

    def check_status(self):
        if int(self.status.item()) != 0:
            raise RuntimeError("fact list contains node/relation ids outside the batch (clamped)")


def csr_build(heads, rels, tails, B, N, R1, nfacts=None):
    """heads/rels/tails: 1-D int64 or int32 CUDA tensors (global rows b*N+local) -> CsrGraph.
    ``nfacts``: optional int32[1] device tensor -- only the first ``nfacts`` slots are facts, the rest is capacity
    padding of fixed-shape buffers (GraphedStep)."""
    heads, rels, tails = _cuda(heads, name="heads"), _cuda(rels, name="rels"), _cuda(tails, name="tails")
    if heads.dtype not in (torch.int64, torch.int32) or rels.dtype != heads.dtype or tails.dtype != heads.dtype:
        raise RuntimeError("fact arrays must share dtype int64 or int32")
    F = heads.numel()
    g = CsrGraph(B, N, F, R1, heads.device)
    L = _L()
    ws_bytes = L.gr_csr_build_workspace_bytes(F, B * N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=heads.device)
    with _OpTimer("csr_build"):
        rc = L.gr_csr_build(_p(heads.contiguous()), _p(rels.contiguous()), _p(tails.contiguous()),
                            heads.element_size(), F, B * N, R1,
                            _p(g.rowptr_t), _p(g.src_t), _p(g.rel_t), _p(g.fact_t),
                            _p(g.rowptr_h), _p(g.src_h), _p(g.rel_h), _p(g.fact_h),
                            _p(g.status), _p(nfacts), _p(ws), ws_bytes, _stream())
    _lib.check(rc)
    STATS.launches += 9 if F > 0 else 5     # memsets excluded: hist, 3x scan, place, 2x sort, fill (+gather)
    return g


def gather_f32(values, fact):
    values = _cuda(values, torch.float32, "values")
    F = values.numel()
    out = torch.empty(max(pad4(F), 4), dtype=torch.float32, device=values.device)
    _lib.check(_L().gr_gather_f32(_p(values), _p(fact), _p(out), F, _stream()))
    STATS.launches += 1
    return out


def linear(A, W, bias=None, relu=False, out=None, addend=None, addend_rows=0, exact=False):
    """out[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ addend rows).  A/out may be strided row views."""
    A, W = _cuda(A, torch.float32, "A"), _cuda(W, torch.float32, "W")
    M, K = A.shape
    N = W.shape[0]
    assert W.shape[1] == K and A.stride(1) == 1 and W.stride(1) == 1
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    flags = (LINEAR_RELU if relu else 0) | (LINEAR_EXACT_FP32 if exact else 0)
    rc = _L().gr_linear(_p(A), A.stride(0), _p(W), W.stride(0), _p(bias), _p(addend),
                        addend.stride(0) if addend is not None else 0, addend_rows,
                        _p(out), out.stride(0), M, N, K, flags, _stream())
    _lib.check(rc)
    STATS.launches += 1
    return out


TC_LINEAR = True       # route the big e2e linears through the tcgen05 split-bf16 kernel (models use this)


def linear_tc(A, W, bias=None, relu=False, out=None):
    """Tensor-core (tcgen05, split-bf16 x3) version of :func:`linear` for 8 <= N <= 256."""
    A, W = _cuda(A, torch.float32, "A"), _cuda(W, torch.float32, "W")
    M, K = A.shape
    N = W.shape[0]
    assert W.shape[1] == K and A.stride(1) == 1 and W.stride(1) == 1
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    L = _L()
    nbytes = L.gr_linear_tc_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    rc = L.gr_linear_tc(_p(A), A.stride(0), _p(W), W.stride(0), _p(bias), _p(out), out.stride(0),
                        M, N, K, LINEAR_RELU if relu else 0, _p(ws), nbytes, _stream())
    _lib.check(rc)
    STATS.launches += 3
    return out


def rel_linear(A, W, bias=None, addend=None, addend_rows=0):
    """Hoisted relation projection table = A W^T + b (+ pos_emb rows): tcgen05 split-bf16 path when enabled
    (fp32-class accuracy, ~3x faster than the SIMT kernel at [6107 x 200 x 200]); the optional pos_emb addend
    keeps the exact SIMT kernel."""
    if TC_LINEAR and addend is None and 8 <= W.shape[0] <= 256 and W.shape[1] >= 8:
        return linear_tc(A, W, bias, relu=False)
    return linear(A, W, bias, addend=addend, addend_rows=addend_rows)


def e2e_linear(A, W, bias, out):
    """relu(A W^T + b) for the node-update GEMM: tcgen05 path when enabled and the shape fits."""
    if TC_LINEAR and 8 <= W.shape[0] <= 256 and W.shape[1] >= 8:
        return linear_tc(A, W, bias, relu=True, out=out)
    return linear(A, W, bias, relu=True, out=out)


def aggregate(g, direction, prior, table, ins, out=None, out_col0=0, seg_stride=None, w=None,
              possible=None):
    """One direction of the relation-typed aggregation.  direction: 'fwd' (tail CSR) | 'inv' (head CSR).
    prior [B,N]; table [R1,D]; ins [B,I,D]; out [B*N, >= out_col0 + I*seg_stride] row-major view."""
    prior = _cuda(prior, torch.float32, "prior").contiguous()
    table = _cuda(table, torch.float32, "table").contiguous()
    ins = _cuda(ins, torch.float32, "ins").contiguous()
    B, I, D = ins.shape
    N = g.N
    if seg_stride is None:
        seg_stride = D
    if out is None:
        out = torch.empty(B * N, I * seg_stride + out_col0, dtype=torch.float32, device=prior.device)
    assert out.stride(1) == 1
    if direction == "fwd":
        rp, src, rel = g.rowptr_t, g.src_t, g.rel_t
    else:
        rp, src, rel = g.rowptr_h, g.src_h, g.rel_h
    with _AggTimer(("single", I)):
        rc = _L().gr_aggregate(_p(rp), _p(src), _p(rel), _p(w), _p(prior), _p(table), _p(ins), _p(out),
                               out.stride(0), out_col0, seg_stride, _p(possible), B, N, D, I, g.F,
                               _stream())
    _lib.check(rc)
    STATS.launches += (I + 3) // 4
    return out


def aggregate_backward(g, direction, prior, table, ins, grad_out, grad_table, grad_ins, grad_prior, w=None):
    """Accumulate the gradients of :func:`aggregate` (same ``direction`` / CSR) into grad_table [R1,D], grad_ins
    [B,I,D], grad_prior [B,N]; grad_out [B*N, I*D] contiguous rows (csrc/aggregate_bwd.cu)."""
    prior = _cuda(prior, torch.float32, "prior").contiguous()
    table = _cuda(table, torch.float32, "table").contiguous()
    ins = _cuda(ins, torch.float32, "ins").contiguous()
    grad_out = _cuda(grad_out, torch.float32, "grad_out")
    B, I, D = ins.shape
    assert grad_out.stride(1) == 1 and grad_table.is_contiguous() and grad_ins.is_contiguous() and grad_prior.is_contiguous()
    rp, src, rel = (g.rowptr_t, g.src_t, g.rel_t) if direction == "fwd" else (g.rowptr_h, g.src_h, g.rel_h)
    with _OpTimer("aggregation_bwd"):
        rc = _L().gr_aggregate_backward(_p(rp), _p(src), _p(rel), _p(w), _p(prior), _p(table), _p(ins), _p(grad_out),
                                        grad_out.stride(0), 0, D, _p(grad_table), _p(grad_ins), _p(grad_prior),
                                        B, g.N, D, I, g.F, _stream())
    _lib.check(rc)
    STATS.launches += 1


def aggregate_dual(g, prior, table_fwd, table_inv, ins, out, out_col0, w_t=None, w_h=None, planes=None,
                   seg_pitch=0):
    """Both directions of one ReaRev layer: out[:, out_col0 + (2j+dir)*D : +D] (reasongnn.py:150-161).
    ``planes`` = (hi, lo) bf16 [B*N, ld] tensors: write the split-bf16 A-operand planes (``out`` may be
    None)."""
    prior = _cuda(prior, torch.float32, "prior").contiguous()
    ins = _cuda(ins, torch.float32, "ins").contiguous()
    B, I, D = ins.shape
    assert table_fwd.is_contiguous() and table_inv.is_contiguous()
    assert out is None or out.stride(1) == 1
    hi, lo = planes if planes is not None else (None, None)
    with _AggTimer(("dual", I)):
        rc = _L().gr_aggregate_dual(_p(g.rowptr_t), _p(g.src_t), _p(g.rel_t), _p(w_t),
                                    _p(g.rowptr_h), _p(g.src_h), _p(g.rel_h), _p(w_h),
                                    _p(prior), _p(table_fwd), _p(table_inv), _p(ins), _p(out),
                                    out.stride(0) if out is not None else 0, out_col0, seg_pitch,
                                    _p(hi), _p(lo), hi.stride(0) if hi is not None else 0,
                                    B, g.N, D, I, g.F, _stream())
    _lib.check(rc)
    STATS.launches += (I + 3) // 4
    return out


AGG_ABS = True          # |v|-accumulating aggregation kernel (csrc/aggregate_abs.cu) for the shapes it specialises


def pad_table256(table):
    """[rows, D] fp32 relation table -> zero-padded [rows, 256] copy (1 KB rows: every lane of the gather is
    in-bounds) for :func:`aggregate_dual_abs`."""
    table = _cuda(table, torch.float32, "table")
    rows, D = table.shape
    assert table.stride(1) == 1
    pn = torch.empty(rows, 256, dtype=torch.float32, device=table.device)
    with _OpTimer("table_prep"):
        _lib.check(_L().gr_pad_table256(_p(table), table.stride(0), rows, D, _p(pn), _stream()))
    STATS.launches += 1
    return pn


def aggregate_dual_abs_supported(N, D, seg_pitch, R1):
    return bool(AGG_ABS) and bool(_L().gr_aggregate_dual_abs_supported(N, D, seg_pitch, R1))


_TILE_COUNTER = {}


def aggregate_dual_abs(g, prior, pn_fwd, pn_inv, ins, planes, out_col0, seg_pitch, w_t=None, w_h=None):
    """Both directions of one ReaRev layer into the split-bf16 planes, |v|-accumulating kernel (reasongnn.py:150-161)."""
    prior = _cuda(prior, torch.float32, "prior").contiguous()
    ins = _cuda(ins, torch.float32, "ins").contiguous()
    B, I, D = ins.shape
    hi, lo = planes
    assert pn_fwd.is_contiguous() and pn_inv.is_contiguous() and hi.stride(0) == lo.stride(0)
    if ACT_BF16:
        lo = None
    dev = prior.device
    if dev not in _TILE_COUNTER:
        _TILE_COUNTER[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    with _AggTimer(("dual", I)):
        rc = _L().gr_aggregate_dual_abs(_p(g.rowptr_t), _p(g.src_t), _p(g.rel_t), _p(w_t),
                                       _p(g.rowptr_h), _p(g.src_h), _p(g.rel_h), _p(w_h),
                                       _p(prior), _p(pn_fwd), _p(pn_inv), pn_fwd.shape[0], _p(ins), _p(hi), _p(lo),
                                       hi.stride(0),
                                       out_col0, seg_pitch, B, g.N, D, I, g.F, _p(_TILE_COUNTER[dev]), _stream())
    _lib.check(rc)
    STATS.launches += (I + 3) // 4


FUSED_LAYER = True      # dense-prior ReaRev layers: aggregation fused into the e2e GEMM (csrc/fused_layer.cu)
FUSED_MIN_ROWS = 148 * 128   # below one 128-row tile per SM the fused kernel's serial per-tile chain (35 dependent k-blocks)
                             # loses to the two wide kernels (cfg1: 0.853 vs 0.836 ms per step)


def fused_layer_supported(N, D, seg_pitch, I, n_out):
    return bool(_L().gr_fused_layer_supported(N, D, seg_pitch, I, n_out))


def fused_ell(g, w_t=None, w_h=None):
    """Quad-ELL form of the batch's CSRs for the fused layer kernel, built once per batch and cached on the graph
    (keyed on whether edge weights are used: they are part of the static entries)."""
    key = "_ell_w" if w_t is not None else "_ell"
    ell = getattr(g, key, None)
    if ell is None:
        L = _L()
        nbytes = L.gr_fused_ell_bytes(g.B, g.N, g.F)
        ell = torch.empty(nbytes, dtype=torch.uint8, device=g.rowptr_t.device)
        with _OpTimer("csr_build"):
            _lib.check(L.gr_fused_ell_build(_p(g.rowptr_t), _p(g.src_t), _p(g.rel_t), _p(w_t),
                                            _p(g.rowptr_h), _p(g.src_h), _p(g.rel_h), _p(w_h),
                                            g.B, g.N, g.F, _p(ell), nbytes, _stream()))
        STATS.launches += 1
        setattr(g, key, ell)
    return ell


def fused_layer(g, prior, pn_fwd, pn_inv, ins, h_planes, seg_pitch, W, bias, out=None, out_planes=None,
                w_score=None, dots=None, relu=True, w_t=None, w_h=None):
    """One dense-prior ReaRev layer in one kernel (reasongnn.py:134-165): both directions and all instructions are
    aggregated straight into the tensor-core operand stages of ``relu(e2e([h | nb...]))``.  ``h_planes`` = (hi, lo)
    bf16 planes whose first ``seg_pitch`` columns hold h; writes any of fp32 ``out``, ``out_planes``, ``dots``."""
    prior = _cuda(prior, torch.float32, "prior").contiguous()
    ins = _cuda(ins, torch.float32, "ins").contiguous()
    B, I, D = ins.shape
    hi, lo = h_planes
    n_out = W.shape[0]
    assert pn_fwd.is_contiguous() and pn_inv.is_contiguous() and hi.stride(0) == lo.stride(0)
    assert W.stride(1) == 1 and W.shape[1] == (2 * I + 1) * D
    L = _L()
    nbytes = L.gr_fused_layer_workspace_bytes(D, seg_pitch, I, n_out)
    ws, presplit = _weight_ws(W, n_out, W.shape[1], "fused", seg_pitch, nbytes)
    ell = fused_ell(g, w_t, w_h)
    chi, clo = out_planes if out_planes is not None else (None, None)
    flags = (LINEAR_RELU if relu else 0) | (LINEAR_W_PRESPLIT if presplit else 0)
    with _OpTimer("fused_layer"):
        rc = L.gr_fused_layer(_p(g.rowptr_t), _p(g.src_t), _p(g.rel_t), _p(w_t),
                              _p(g.rowptr_h), _p(g.src_h), _p(g.rel_h), _p(w_h),
                              _p(prior), _p(pn_fwd), _p(pn_inv), _p(ins), _p(hi), _p(lo), hi.stride(0), seg_pitch,
                              _p(W), W.stride(0), _p(bias), _p(out), out.stride(0) if out is not None else 0,
                              _p(chi), _p(clo), chi.stride(0) if chi is not None else 0, _p(w_score), _p(dots),
                              B, g.N, D, I, n_out, g.F, flags, _p(ws), ws.numel(), _p(ell), ell.numel(), _stream())
    _lib.check(rc)
    STATS.launches += 2 if (w_t is not None or w_h is not None) else 1     # weighted graphs: + the coefficient pass
    return out


def type_layer(g, table, out, w_t=None, w_h=None, planes=None):
    """out[:, :D] = relu(sum_tail w*table[rel] + sum_head w*table[rel]) (layer_init.py:46-57); optional
    split-bf16 planes of the same values."""
    table = _cuda(table, torch.float32, "table").contiguous()
    D = table.shape[1]
    assert out is None or out.stride(1) == 1
    hi, lo = planes if planes is not None else (None, None)
    with _OpTimer("type_layer"):
        rc = _L().gr_type_layer(_p(g.rowptr_t), _p(g.rel_t), _p(w_t), _p(g.rowptr_h), _p(g.rel_h), _p(w_h),
                                _p(table), _p(out), out.stride(0) if out is not None else 0,
                                _p(hi), _p(lo), hi.stride(0) if hi is not None else 0,
                                g.B, g.N, D, g.F, _stream())
    _lib.check(rc)
    STATS.launches += 1
    return out


def split_bf16(A, hi, lo):
    """fp32 [M,K] -> bf16 hi/lo planes (first K columns of hi/lo)."""
    A = _cuda(A, torch.float32, "A")
    M, K = A.shape
    assert A.stride(1) == 1 and hi.stride(1) == 1 and hi.stride(0) == lo.stride(0)
    _lib.check(_L().gr_split_bf16(_p(A), A.stride(0), M, K, _p(hi), _p(lo), hi.stride(0), _stream()))
    STATS.launches += 1


WEIGHT_CACHE = True    # keep the bf16 hi/lo split of a weight matrix until its tensor version changes
_W_CACHE = {}          # (ptr, ldw, N, K, k_seg, pitch) -> [workspace, version, weakref(base tensor)]
_P_CACHE = {}          # ptr -> [hi, lo, version, weakref(tensor)]


def _base(t):
    return t._base if t._base is not None else t


def _weight_ws(W, N, K, k_seg, k_seg_pitch, nbytes):
    """Workspace holding W's split planes + whether it is still valid (weight pre-formatting: the conversion
    runs once per weight VERSION, so an in-place update / load_state_dict re-splits on the next call)."""
    capturing = torch.cuda.is_current_stream_capturing()
    key = (W.data_ptr(), W.stride(0), N, K, k_seg, k_seg_pitch)
    ent = _W_CACHE.get(key) if WEIGHT_CACHE else None
    if ent is not None and ent[2]() is _base(W) and ent[0].numel() >= nbytes:
        if ent[1] == W._version:
            return ent[0], True
        if not capturing:
            ent[1] = W._version
            return ent[0], False          # same buffer, re-split
    ws = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    if WEIGHT_CACHE and not capturing:
        if len(_W_CACHE) > 512:          # models that were dropped: release the workspaces of dead weights
            for k in [k for k, e in _W_CACHE.items() if e[2]() is None]:
                del _W_CACHE[k]
        _W_CACHE[key] = [ws, W._version, weakref.ref(_base(W))]
    return ws, False


def param_planes(P):
    """bf16 hi/lo planes [M, round_up(K, 64)] of a parameter matrix used as a GEMM A operand (relation embedding
    tables), cached per tensor version like the weight split."""
    M, K = P.shape
    capturing = torch.cuda.is_current_stream_capturing()
    ent = _P_CACHE.get(P.data_ptr()) if WEIGHT_CACHE else None
    if ent is not None and ent[3]() is P and ent[0].shape[0] == M and ent[2] == P._version:
        return ent[0], ent[1]
    if ent is not None and ent[3]() is P and ent[0].shape[0] == M and not capturing:
        hi, lo = ent[0], ent[1]
        ent[2] = P._version
    else:
        Kp = (K + 63) // 64 * 64
        hi = torch.zeros(M, Kp, dtype=torch.bfloat16, device=P.device)
        lo = torch.zeros(M, Kp, dtype=torch.bfloat16, device=P.device)
        if WEIGHT_CACHE and not capturing:
            _P_CACHE[P.data_ptr()] = [hi, lo, P._version, weakref.ref(P)]
    split_bf16(P.detach(), hi, lo)
    return hi, lo


def clear_weight_cache():
    """Drop the cached pre-formatted weights.  Captured CUDA graphs keep their own references to the workspaces they
    read (:func:`live_weight_workspaces`), so clearing the cache never frees memory a graph replay still uses.
    Note: the caches are validated by ``tensor._version``; in-place writes through ``.data`` (``p.data.copy_``,
    ``p.data.mul_``) do not bump it -- call this function after such updates."""
    _W_CACHE.clear()
    _P_CACHE.clear()


def live_weight_workspaces():
    """Strong references to every cached pre-formatted weight buffer (held by GraphedStep entries)."""
    return [e[0] for e in _W_CACHE.values()] + [t for e in _P_CACHE.values() for t in e[:2]]


TC_MAX_N = 256         # output columns of one tcgen05 GEMM launch (one TMEM accumulator buffer)
TC_MAX_N_SPLIT = 512   # wider outputs are tiled over N: one launch per <= 256-column slice of W


def linear_tc_planes(a_hi, a_lo, K, W, bias, out=None, out_planes=None, w_score=None, dots=None, relu=True,
                     k_seg=0, k_seg_pitch=0, single_ok=False):
    """tcgen05 split-bf16 GEMM whose A operand already lives in bf16 hi/lo planes [M, >=K].
    ``single_ok``: this call may run as ONE bf16 product when ``ACT_BF16`` is on (the node-update GEMMs; the small
    relation-table GEMMs always keep the three-product fp32-class path).
    Writes any of: fp32 ``out`` [M,N]; ``out_planes`` (hi, lo) [M, >=N] (next layer's h columns);
    ``dots`` [2*M] = the two column-half partial sums of out @ w_score.
    N > 256 (cfg5: entity_dim 400) is tiled over the output columns: one launch per slice of W rows."""
    N = W.shape[0]
    if N > TC_MAX_N:
        nsl = (N + TC_MAX_N - 1) // TC_MAX_N
        step = ((N + nsl - 1) // nsl + 15) // 16 * 16
        part = None
        for n0 in range(0, N, step):
            n1 = min(N, n0 + step)
            d = torch.empty_like(dots) if (dots is not None and n0 > 0) else dots
            linear_tc_planes(a_hi, a_lo, K, W[n0:n1], None if bias is None else bias[n0:n1],
                             out=None if out is None else out[:, n0:n1],
                             out_planes=None if out_planes is None else (out_planes[0][:, n0:n1], out_planes[1][:, n0:n1]),
                             w_score=None if w_score is None else w_score[n0:n1], dots=d, relu=relu, k_seg=k_seg,
                             k_seg_pitch=k_seg_pitch, single_ok=single_ok)
            if dots is not None and n0 > 0:
                part = d if part is None else part + d
        if part is not None:
            dots += part
        return out
    M = a_hi.shape[0]
    assert a_hi.dtype == torch.bfloat16 and a_hi.stride(1) == 1 and a_hi.stride(0) == a_lo.stride(0)
    if k_seg and k_seg_pitch > k_seg:
        assert K % k_seg_pitch == 0 and W.shape[1] == K // k_seg_pitch * k_seg
    else:
        assert W.shape[1] == K
    assert W.stride(1) == 1
    L = _L()
    nbytes = L.gr_linear_tc_planes_workspace_bytes(N, K)
    ws, presplit = _weight_ws(W, N, K, k_seg, k_seg_pitch, nbytes)
    chi, clo = out_planes if out_planes is not None else (None, None)
    flags = (LINEAR_RELU if relu else 0) | (LINEAR_W_PRESPLIT if presplit else 0) | \
        (LINEAR_BF16_SINGLE if (ACT_BF16 and single_ok) else 0)
    with _OpTimer("gemm_tc", (M, N, K)):
        rc = L.gr_linear_tc_planes(_p(a_hi), _p(a_lo), a_hi.stride(0), _p(W), W.stride(0), _p(bias),
                                   _p(out), out.stride(0) if out is not None else 0,
                                   _p(chi), _p(clo), chi.stride(0) if chi is not None else 0,
                                   _p(w_score), _p(dots), M, N, K, k_seg, k_seg_pitch,
                                   flags, _p(ws), nbytes, _stream())
    _lib.check(rc)
    STATS.launches += 1 if presplit else 2
    return out


class RelFeatures:
    """Relation features (ReaRev.get_rel_feature / NSM.get_rel_feature) for all directions, stacked row-wise
    [n_dir * R1, D]: as split-bf16 planes (A operand of the per-layer relation-table GEMMs) and/or fp32."""

    _buf = {}

    def __init__(self, R1, D, n_dir, device, planes):
        self.R1, self.D, self.n_dir = R1, D, n_dir
        self.hi = self.lo = self.f32 = None
        if planes:
            Kp = (D + 63) // 64 * 64
            key = (str(device), n_dir * R1, Kp)
            if key not in RelFeatures._buf:
                RelFeatures._buf[key] = (torch.zeros(n_dir * R1, Kp, dtype=torch.bfloat16, device=device),
                                         torch.zeros(n_dir * R1, Kp, dtype=torch.bfloat16, device=device))
            self.hi, self.lo = RelFeatures._buf[key]
        else:
            self.f32 = torch.empty(n_dir * R1, D, dtype=torch.float32, device=device)

    def rows(self, d):
        return slice(d * self.R1, (d + 1) * self.R1)


def _tc_ok(n_out, k_in):
    return bool(TC_LINEAR) and 8 <= n_out <= TC_MAX_N_SPLIT and k_in >= 8


def rel_features_from_embeddings(embs, W, bias):
    """relation_linear applied to the relation embedding table(s) (rearev.py:91-99 / nsm.py:97-104):
    one tcgen05 GEMM per direction straight into the stacked planes (no fp32 round trip)."""
    R1, K = embs[0].shape
    D = W.shape[0]
    planes = _tc_ok(D, K) and _tc_ok(D, D)
    rf = RelFeatures(R1, D, len(embs), W.device, planes)
    for d, E in enumerate(embs):
        if planes:
            ahi, alo = param_planes(E)
            linear_tc_planes(ahi, alo, K, W, bias, out_planes=(rf.hi[rf.rows(d)], rf.lo[rf.rows(d)]), relu=False)
        else:
            linear(E, W, bias, out=rf.f32[rf.rows(d)])
    return rf


def rel_features_from_tensors(feats):
    """Relation features computed elsewhere in fp32 (relation-text encoder, rearev.py:100-111)."""
    R1, D = feats[0].shape
    planes = _tc_ok(D, D)
    rf = RelFeatures(R1, D, len(feats), feats[0].device, planes)
    for d, f in enumerate(feats):
        if planes:
            split_bf16(f.contiguous(), rf.hi[rf.rows(d)], rf.lo[rf.rows(d)])
        else:
            rf.f32[rf.rows(d)].copy_(f)
    return rf


def rel_table(rf, W, bias, dirs=None, addends=None):
    """Hoisted relation projection table(s) = rel_features W^T + b for the stacked directions in ONE GEMM
    (reasongnn.py:79,105 / nsm_gnn.py:95 / layer_init.py:41 applied to R1 relation rows instead of F facts).
    Returns the list of per-direction [R1, D] tables.  ``addends``: optional per-direction pos_emb rows."""
    n = rf.n_dir if dirs is None else dirs
    rows = n * rf.R1
    out = torch.empty(rows, W.shape[0], dtype=torch.float32, device=W.device)
    if rf.hi is not None:
        linear_tc_planes(rf.hi[:rows], rf.lo[:rows], rf.D, W, bias, out=out, relu=False)
    else:
        linear(rf.f32[:rows], W, bias, out=out)
    tabs = [out[d * rf.R1:(d + 1) * rf.R1] for d in range(n)]
    if addends is not None:
        for t, a in zip(tabs, addends):
            if a is not None:
                t[: a.shape[0]] += a
    return tabs


SPARSE_PRIOR_FASTPATH = True   # first layer of every ReaRev iteration (seed prior): K=1-segment GEMM + frontier fix-up


def frontier_rows(g, prior, rows, count):
    """rows/count <- destination rows with at least one in-edge (either direction) from a node with prior != 0."""
    prior = _cuda(prior, torch.float32, "prior").contiguous()
    with _OpTimer("frontier"):
        _lib.check(_L().gr_frontier_rows(_p(g.rowptr_t), _p(g.src_t), _p(g.rowptr_h), _p(g.src_h), _p(prior),
                                         g.B * g.N, _p(rows), _p(count), _stream()))
    STATS.launches += 1


def frontier_fixup(g, prior, table_fwd, table_inv, ins, cur_planes, W, bias, w_score, nxt_planes, h32, dots,
                   rows, count, w_t=None, w_h=None):
    """Recompute the listed rows of one ReaRev layer in full (aggregation + e2e linear + relu + score dot)."""
    prior = _cuda(prior, torch.float32, "prior").contiguous()
    ins = _cuda(ins, torch.float32, "ins").contiguous()
    B, I, D = ins.shape
    chi, clo = cur_planes
    nhi, nlo = nxt_planes
    assert W.stride(1) == 1 and table_fwd.is_contiguous() and table_inv.is_contiguous()
    with _OpTimer("frontier"):
        rc = _L().gr_frontier_fixup(_p(g.rowptr_t), _p(g.src_t), _p(g.rel_t), _p(w_t), _p(g.rowptr_h), _p(g.src_h),
                                    _p(g.rel_h), _p(w_h), _p(prior), _p(table_fwd), _p(table_inv), _p(ins),
                                    _p(chi), _p(clo), chi.stride(0), _p(W), W.stride(0), _p(bias), _p(w_score),
                                    _p(nhi), _p(nlo), nhi.stride(0), _p(h32), _p(dots), _p(rows), _p(count),
                                    B, g.N, D, I, _stream())
    _lib.check(rc)
    STATS.launches += 1


def masked_softmax(dots, b_score, mask, B, N):
    """dist[b,:] = softmax(dots[0,b,:] + dots[1,b,:] + b + (1-mask)*VERY_NEG)  (reasongnn.py:168-169);
    ``dots`` = the [2, B*N] partial score dots of :func:`linear_tc_planes`."""
    dist = torch.empty(B, N, dtype=torch.float32, device=dots.device)
    d = dots.view(2, -1)
    with _OpTimer("softmax"):
        _lib.check(_L().gr_masked_softmax(_p(d[0]), _p(d[1]), _p(b_score), _p(mask.contiguous()), _p(dist), B, N,
                                          _stream()))
    STATS.launches += 1
    return dist


def score_softmax(h, w_score, b_score, mask, B, N, logits_out=None):
    """h: [B*N, >=D] row view (stride(0) = ldh); returns dist [B,N]."""
    h = _cuda(h, torch.float32, "h")
    D = w_score.numel()
    assert h.stride(1) == 1
    dist = torch.empty(B, N, dtype=torch.float32, device=h.device)
    rc = _L().gr_score_softmax(_p(h), h.stride(0), _p(w_score.contiguous()), _p(b_score),
                               _p(mask.contiguous()), _p(dist), _p(logits_out), B, N, D, _stream())
    _lib.check(rc)
    STATS.launches += 2
    return dist


def seed_retrieve(seed_info, h, B, N, D):
    seed_info = _cuda(seed_info, torch.float32, "seed_info").contiguous()
    out = torch.empty(B, D, dtype=torch.float32, device=h.device)
    assert h.stride(1) == 1
    _lib.check(_L().gr_seed_retrieve(_p(seed_info), _p(h), h.stride(0), _p(out), B, N, D, _stream()))
    STATS.launches += 1
    return out


def _ptr_array(tensors):
    for t in tensors:
        _cuda(t, torch.float32, "weight")
        assert t.is_contiguous()
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def instructions(hidden, qnode, qtext, pad_id, Wq, bq, Wcq, bcq, wca, bca):
    """All ``num_ins`` instruction vectors of one question batch in one launch (base_encoder.py:73-114).
    hidden [B,Q,D], qnode [B,D], qtext int64 [B,Q]; Wq/bq: lists of question_linear_i weight/bias.
    Returns ins [B, I, D]."""
    hidden = _cuda(hidden, torch.float32, "hidden").contiguous()
    qnode = _cuda(qnode, torch.float32, "qnode").contiguous()
    qtext = _cuda(qtext, torch.int64, "qtext").contiguous()
    B, Q, D = hidden.shape
    I = len(Wq)
    out = torch.empty(B, I, D, dtype=torch.float32, device=hidden.device)
    with _OpTimer("question_side"):
        rc = _L().gr_instructions(_p(hidden), _p(qnode), _p(qtext), int(pad_id), _ptr_array(Wq), _ptr_array(bq),
                                  _p(Wcq.contiguous()), _p(bcq), _p(wca.contiguous()), _p(bca), _p(out), None,
                                  B, Q, D, I, _stream())
    _lib.check(rc)
    STATS.launches += 1
    return out


LSTM_MAX_HIDDEN = 256


def lstm_forward(gates_x, W_hh, b_hh):
    """hidden [B,Q,D] of a one-layer LSTM with zero initial state, given gates_x = x W_ih^T + b_ih [B,Q,4D]
    (lstm_encoder.py:27-36); one launch for the whole sequence."""
    gates_x = _cuda(gates_x, torch.float32, "gates_x").contiguous()
    B, Q, G = gates_x.shape
    D = G // 4
    assert W_hh.shape == (4 * D, D) and W_hh.is_contiguous()
    hidden = torch.empty(B, Q, D, dtype=torch.float32, device=gates_x.device)
    with _OpTimer("question_side"):
        _lib.check(_L().gr_lstm_forward(_p(gates_x), _p(W_hh), _p(b_hh), _p(hidden), B, Q, D, _stream()))
    STATS.launches += 1
    return hidden


def query_reform(seed_info, h, ins, Wr, Wg, B, N):
    """ins_new[b,j] = Fusion_j(ins[b,j], seed_info[b] @ h[b]) for every instruction (query_update.py:6-44)."""
    seed_info = _cuda(seed_info, torch.float32, "seed_info").contiguous()
    ins = _cuda(ins, torch.float32, "ins").contiguous()
    h = _cuda(h, torch.float32, "h")
    assert h.stride(1) == 1
    _, I, D = ins.shape
    out = torch.empty_like(ins)
    with _OpTimer("query_reform"):
        rc = _L().gr_query_reform(_p(seed_info), _p(h), h.stride(0), _p(ins), _ptr_array(Wr), _ptr_array(Wg),
                                  _p(out), None, B, N, D, I, _stream())
    _lib.check(rc)
    STATS.launches += 1
    return out


def kl_loss_pred(dist, teacher):
    """-> (loss 0-dim fp32, pred int64[B]) : calc_loss_label('kl') with case_valid, and argmax (base_model.py:186)."""
    dist = _cuda(dist, torch.float32, "dist").contiguous()
    teacher = _cuda(teacher, torch.float32, "teacher").contiguous()
    B, N = dist.shape
    loss_q = torch.empty(B, dtype=torch.float32, device=dist.device)
    loss = torch.empty((), dtype=torch.float32, device=dist.device)
    pred = torch.empty(B, dtype=torch.int64, device=dist.device)
    with _OpTimer("loss_rank"):
        _lib.check(_L().gr_kl_loss_pred(_p(dist), _p(teacher), _p(loss_q), _p(loss), _p(pred), B, N, _stream()))
    STATS.launches += 2
    return loss, pred


def rank_candidates(dist, local_entity, query_entities, pad_id, eps):
    """-> (cand_idx int32[B,N], cand_count int32[B], cand_total int32[B]) on device."""
    dist = _cuda(dist, torch.float32, "dist").contiguous()
    local_entity = _cuda(local_entity, torch.int64, "local_entity").contiguous()
    query_entities = _cuda(query_entities, torch.float32, "query_entities").contiguous()
    B, N = dist.shape
    dev = dist.device
    cand_idx = torch.empty(B, N, dtype=torch.int32, device=dev)
    cand_count = torch.empty(B, dtype=torch.int32, device=dev)
    cand_total = torch.empty(B, dtype=torch.int32, device=dev)
    L = _L()
    nbytes = L.gr_rank_workspace_bytes(B, N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with _OpTimer("loss_rank"):
        rc = L.gr_rank_candidates(_p(dist), _p(local_entity), _p(query_entities), int(pad_id), float(eps),
                                  _p(cand_idx), _p(cand_count), _p(cand_total), B, N, _p(ws), nbytes,
                                  _stream())
    _lib.check(rc)
    STATS.launches += 1
    return cand_idx, cand_count, cand_total


def shortest_path_nodes(g, source_idx, source_cnt, target_idx, target_cnt, return_distances=False):
    """source_idx int32[B,S], target_idx int32[B,T] local indices (+counts) ->
    (on_path uint8[B,N], pair_dist int32[B,S,T]); with ``return_distances`` also the BFS distance arrays the kernel
    leaves in its workspace: int32 [B, S+T, N] (sources first), -1 = unreachable."""
    B, N = g.B, g.N
    S, T = source_idx.shape[1], target_idx.shape[1]
    dev = source_idx.device
    on_path = torch.empty(B, N, dtype=torch.uint8, device=dev)
    pair_dist = torch.empty(B, S, T, dtype=torch.int32, device=dev)
    L = _L()
    nbytes = L.gr_paths_workspace_bytes(B, N, S, T)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = L.gr_shortest_path_nodes(_p(g.rowptr_t), _p(g.src_t), _p(g.rowptr_h), _p(g.src_h),
                                  _p(source_idx.contiguous()), _p(source_cnt.contiguous()), S,
                                  _p(target_idx.contiguous()), _p(target_cnt.contiguous()), T,
                                  _p(on_path), _p(pair_dist), B, N, _p(ws), nbytes, _stream())
    _lib.check(rc)
    if return_distances:
        return on_path, pair_dist, ws[: B * (S + T) * N * 4].view(torch.int32).view(B, S + T, N)
    return on_path, pair_dist
