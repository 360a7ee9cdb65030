"""Differentiable forward for ``model(batch, training=True)`` (gnn/train_model.py:209-233).

The inference path runs hand-written CUDA kernels without a backward; ``Trainer_KBQA.train_epoch`` needs
``loss.backward()`` through the same parameters.  This module evaluates the same math with torch ops on the model's
device (CUDA when the model lives there -- no CPU fallback is involved) so that autograd provides the gradients:

  * the relation projection ``rel_linear_k`` is hoisted from the F facts to the R1 relation rows (one GEMM per layer);
  * messages are formed per fact (``relu(P[rel] * ins[batch]) * w^2 * prior[src]``) and reduced with ``index_add_``
    (deterministic order is not needed for training);
  * dropouts are applied where the reference applies them (``linear_drop`` before e2e / score / instruction linears,
    ``lstm_drop`` on the word embeddings), active only under ``model.train()``.

Reference: ReaRev.forward gnn/models/ReaRev/rearev.py:163-243, ReasonGNNLayer.forward gnn/modules/kg_reasoning/
reasongnn.py:61-174, TypeLayer.forward gnn/modules/layer_init.py:25-62, BaseInstruction.get_instruction
gnn/modules/question_encoding/base_encoder.py:82-102, QueryReform / Fusion gnn/modules/query_update.py:6-44,
NSM.forward gnn/models/NSM/nsm.py:179-254, NSMLayer gnn/modules/kg_reasoning/nsm_gnn.py:54-112, the loss
gnn/models/base_model.py:193-215 and the train-time metrics get_eval_metric gnn/models/base_model.py:236-298.
"""
import numpy as np
import torch
import torch.nn.functional as F

VERY_NEG_NUMBER = -100000000000


class _Facts:
    """Fact arrays of one ``get_batch`` tuple on the device (the COO operators of build_matrix, base_gnn.py:19-51,
    as index vectors)."""

    def __init__(self, kb_adj_mat, device, normalized_gnn, norm_rel):
        heads, rels, tails, bids, _fids, weight_list, weight_rel_list = kb_adj_mat

        def idx(a):
            if isinstance(a, torch.Tensor):
                return a.to(device=device, dtype=torch.int64)
            return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.int64))).to(device)
        self.heads, self.rels, self.tails = idx(heads), idx(rels), idx(tails)
        if bids is None:
            raise ValueError("training needs the batch_ids array of kb_adj_mat (dataset_load.py:521)")
        self.bids = idx(bids)
        self.w = self.wr = None
        if normalized_gnn:
            if weight_list is None:
                raise ValueError("normalized_gnn needs kb_adj_mat's weight_list")
            self.w = torch.as_tensor(np.asarray(weight_list, dtype=np.float32), device=device)
        if norm_rel:
            if weight_rel_list is None:
                raise ValueError("norm_rel needs kb_adj_mat's weight_rel_list")
            self.wr = torch.as_tensor(np.asarray(weight_rel_list, dtype=np.float32), device=device)


def _scatter_rows(values, dst, rows):
    out = torch.zeros(rows, values.shape[1], dtype=values.dtype, device=values.device)
    return out.index_add_(0, dst, values)


def _type_layer(layer, facts, rel_features, Nt):
    """layer_init.py:44-59: relu(sum over facts into tails + sum over facts into heads) of kb_self_linear(rel)."""
    fact_val = layer.kb_self_linear(rel_features)[facts.rels]
    if facts.wr is not None:
        fact_val = fact_val * facts.wr.unsqueeze(1)
    return F.relu(_scatter_rows(fact_val, facts.tails, Nt) + _scatter_rows(fact_val, facts.heads, Nt))


def _aggregate(table, ins_j, prior_flat, facts, src, dst, Nt):
    """reasongnn.py:61-89 (src = heads, dst = tails) / :91-116 (src = tails, dst = heads)."""
    fact_val = F.relu(table[facts.rels] * ins_j[facts.bids])
    prior = prior_flat[src]
    if facts.w is not None:
        prior = prior * facts.w * facts.w          # head2fact and fact2tail both carry the weight (base_gnn.py:38-48)
    return _scatter_rows(fact_val * prior.unsqueeze(1), dst, Nt)


USE_KERNELS = True      # CUDA tensors: aggregation forward / backward through the hand-written kernels (below)
HOST_CHECK = False      # tests only: let a CPU-resident model evaluate this restatement with torch CPU ops, so that the
                        # ``-m "not gpu"`` suite can hold it against the reference's gradients.  Off (the product):
                        # ``model(batch, training=True)`` on a CPU model raises, like the inference path.


def _require_cuda(dev):
    if dev.type != "cuda" and not HOST_CHECK:
        raise RuntimeError("gnn_rag_b200 runs on CUDA only: move the model to a B200 (model.cuda()); there is no CPU "
                           "path (training=True included)")


class _AggregateFn(torch.autograd.Function):
    """out[n, j, :] = sum_{e -> n} w_e^2 p[src_e] relu(table[rel_e] * ins[b, j]) for all instructions j of one direction:
    forward = gr_aggregate (csrc/aggregate.cu), backward = gr_aggregate_backward (csrc/aggregate_bwd.cu)."""

    @staticmethod
    def forward(ctx, table, ins, prior, graph, direction, w):
        from . import ops
        out = ops.aggregate(graph, direction, prior.detach(), table.detach(), ins.detach(), w=w)
        ctx.save_for_backward(table, ins, prior)
        ctx.graph, ctx.direction, ctx.w = graph, direction, w
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from . import ops
        table, ins, prior = ctx.saved_tensors
        gt, gi, gp = torch.zeros_like(table), torch.zeros_like(ins), torch.zeros_like(prior)
        ops.aggregate_backward(ctx.graph, ctx.direction, prior, table.contiguous(), ins.contiguous(),
                               grad_out.contiguous(), gt, gi, gp, ctx.w)
        return gt, gi, gp, None, None, None


def _kernel_graph(model, batch, device, D, I):
    """CSR of the batch for the kernel path, or None (CPU tensors / shapes the backward kernel does not cover)."""
    if not (USE_KERNELS and device.type == "cuda" and D <= 256 and I <= 4):
        return None
    from . import batching
    return batching.stage_batch(batch, device, model.num_relation + 1, model.normalized_gnn, False).graph


def _neighbours(table_f, table_i, ins, dist, facts, graph, Nt):
    """[Nt, I, n_dir, D] neighbour messages of one layer (reasongnn.py:150-156), kernel or torch path."""
    I = ins.shape[1]
    if graph is not None:
        outs = [_AggregateFn.apply(table_f.contiguous(), ins, dist, graph, "fwd", graph.w_t).view(Nt, I, 1, -1)]
        if table_i is not None:
            outs.append(_AggregateFn.apply(table_i.contiguous(), ins, dist, graph, "inv", graph.w_h).view(Nt, I, 1, -1))
        return torch.cat(outs, dim=2)
    pf = dist.reshape(-1)
    reps = []
    for j in range(I):
        r = [_aggregate(table_f, ins[:, j], pf, facts, facts.heads, facts.tails, Nt)]
        if table_i is not None:
            r.append(_aggregate(table_i, ins[:, j], pf, facts, facts.tails, facts.heads, Nt))
        reps.append(torch.stack(r, dim=1))
    return torch.stack(reps, dim=1)


def _instructions(enc, q_input):
    """base_encoder.py:73-114 on top of encode_question; returns [B, num_ins, D]."""
    enc.encode_question_train(q_input)
    hidden, qnode, qmask = enc.query_hidden_emb, enc.query_node_emb, enc.query_mask_train
    drop = enc.linear_drop
    rel_ins = torch.zeros(q_input.size(0), enc.entity_dim, device=q_input.device)
    out = []
    for i in range(enc.num_ins):
        ri = rel_ins.unsqueeze(1)
        q_i = getattr(enc, "question_linear" + str(i))(drop(qnode))
        cq = enc.cq_linear(drop(torch.cat((ri, q_i, q_i - ri, q_i * ri), dim=-1)))
        ca = enc.ca_linear(drop(cq * hidden))
        attn = F.softmax(ca + (1 - qmask.unsqueeze(2)) * VERY_NEG_NUMBER, dim=1)
        rel_ins = torch.sum(attn * hidden, dim=1)
        out.append(rel_ins)
    return torch.stack(out, dim=1)


def _loss(model, pred_dist, answer_dist):
    case_valid = (torch.sum(answer_dist, dim=1, keepdim=True) > 0).float()
    return model.calc_loss_label(pred_dist, answer_dist, case_valid)


def _retrieved_sets_host(pred_dist, local_entity, seeds, num_entity, eps):
    """Candidate cut of f1_and_hits (base_model.py:216-234) with torch ops on the tensors' own device: used when the
    training tensors do not live on a GPU (the CUDA path uses the ranking kernel)."""
    B, N = pred_dist.shape
    keep = (seeds == 0) & (local_entity != num_entity) & (pred_dist >= (1 - eps) / N)
    p = torch.where(keep, pred_dist, torch.full_like(pred_dist, -1.0))
    order = torch.sort(p, dim=1, descending=True, stable=True)[1]
    ps = torch.gather(p, 1, order)
    csum = torch.cumsum(torch.where(ps >= 0, ps, torch.zeros_like(ps)).double(), dim=1)
    total = keep.sum(1)
    crossed = (csum > eps) & (ps >= 0)
    first = torch.where(crossed.any(1), crossed.float().argmax(1) + 1, total)
    count = torch.minimum(first, total)
    return order, count


@torch.no_grad()
def eval_metric(model, pred_dist, answer_dist, seed_dist, local_entity):
    """get_eval_metric (base_model.py:281-298): hit@1 per question, and F1 of the eps-mass retrieval for the questions
    that have hit@1 (0 for the others, :250-253).  The retrieval runs in the ranking kernel when the tensors are on the
    GPU; answers = non-seed, non-pad candidates with answer mass (:266-271)."""
    top1 = pred_dist.argmax(dim=-1, keepdim=True)
    h1 = ((torch.zeros_like(pred_dist).scatter_(1, top1, 1.0) * (answer_dist > 1e-10).float()).sum(-1) > 0).float()
    f1 = torch.zeros_like(h1)
    if bool(h1.any()):
        seeds = (seed_dist > 0).float()
        if pred_dist.is_cuda:
            from . import ops
            cand_idx, cand_count, _ = ops.rank_candidates(pred_dist.contiguous(), local_entity, seeds,
                                                          model.num_entity, model.eps)
        else:
            cand_idx, cand_count = _retrieved_sets_host(pred_dist, local_entity, seeds, model.num_entity, model.eps)
        counts = cand_count.cpu().tolist()
        idx_h = cand_idx.cpu().numpy()
        ans_ok = ((answer_dist > 0) & (seeds == 0) & (local_entity != model.num_entity)).cpu().numpy()
        le_h = local_entity.cpu().numpy()
        h1_h = h1.cpu().tolist()
        vals = []
        for b, c in enumerate(counts):
            if h1_h[b] == 0.0:
                vals.append(0.0)
                continue
            ans_ids = le_h[b][ans_ok[b]]                  # answers are ENTITY ids (a list: :266-271), as are candidates
            n_ans = len(ans_ids)
            correct = int(np.isin(le_h[b][idx_h[b, :c]], ans_ids).sum()) if c else 0
            if n_ans == 0:
                vals.append(1.0 if c == 0 else 0.0)
            elif c == 0 or correct == 0:
                vals.append(0.0)
            else:
                p, r = correct / c, correct / n_ans
                vals.append(2.0 / (1.0 / p + 1.0 / r))
        f1 = torch.tensor(vals, dtype=torch.float32, device=pred_dist.device)
    return h1, f1


def _stage(model, batch):
    local_entity, query_entities, kb_adj_mat, q_input, seed_dist, _tb, answer_dist = batch[:7]
    dev = model.word_embedding.weight.device
    _require_cuda(dev)

    def t(x, dtype):
        x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return x.to(device=dev, dtype=dtype)
    facts = _Facts(kb_adj_mat, dev, model.normalized_gnn, model.norm_rel)
    return (t(local_entity, torch.int64), t(query_entities, torch.float32), facts, t(q_input, torch.int64),
            t(seed_dist, torch.float32), t(answer_dist, torch.float32))


def rearev_forward(model, batch):
    """ReaRev forward with autograd (rearev.py:163-243) -> (loss, pred, pred_dist, [h1, f1])."""
    local_entity, query_entities, facts, q_input, seed_dist, answer_dist = _stage(model, batch)
    B, N = local_entity.shape
    Nt, D, I = B * N, model.entity_dim, model.num_ins
    layer = model.reasoning
    rel_f, rel_f_inv = model.get_rel_feature_train()
    if model.encode_type:
        h = _type_layer(model.type_layer, facts, rel_f, Nt)
    else:
        h = model.entity_linear(model.entity_embedding(local_entity)).view(Nt, D)
    instructions = _instructions(model.instruction, q_input)           # [B, I, D]
    ins_list = [instructions[:, j] for j in range(I)]
    mask = (local_entity != model.num_entity).float()
    drop = layer.linear_drop_train
    tables = []
    for k in range(model.num_gnn):
        lin = getattr(layer, "rel_linear" + str(k))
        tf, ti = lin(rel_f), lin(rel_f_inv)
        if layer.use_posemb:
            pe, pei = getattr(layer, "pos_emb" + str(k)).weight, getattr(layer, "pos_emb_inv" + str(k)).weight
            tf = torch.cat([tf[: pe.shape[0]] + pe, tf[pe.shape[0]:]])
            ti = torch.cat([ti[: pei.shape[0]] + pei, ti[pei.shape[0]:]])
        tables.append((tf, ti))
    dist_history = [seed_dist]
    dist = seed_dist
    graph = _kernel_graph(model, batch, h.device, D, I)
    for _t in range(model.num_iter):
        dist = seed_dist
        ins = torch.stack(ins_list, dim=1)                                  # [B, I, D]
        for k in range(model.num_gnn):
            tf, ti = tables[k]
            nb = _neighbours(tf, ti, ins, dist, facts, graph, Nt)           # [Nt, I, 2, D]: (j, direction) as in :150-156
            h = F.relu(getattr(layer, "e2e_linear" + str(k))(drop(torch.cat([h, nb.reshape(Nt, -1)], dim=1))))
            score = layer.score_func(drop(h)).view(B, N) + (1 - mask) * VERY_NEG_NUMBER
            dist = F.softmax(score, dim=1)
        dist_history.append(dist)
        hB = h.view(B, N, D)
        new = []
        for j in range(I):
            reform = getattr(model, "reform" + str(j))
            seed_retrieve = torch.bmm(query_entities.unsqueeze(1), hB).squeeze(1)
            new.append(reform.fusion(ins_list[j], seed_retrieve))
        ins_list = new
    pred_dist = dist_history[-1]
    loss = _loss(model, pred_dist, answer_dist)
    pred = torch.max(pred_dist, dim=1)[1]
    h1, f1 = eval_metric(model, pred_dist.detach(), answer_dist, seed_dist, local_entity)
    model.dist_history = dist_history
    return loss, pred, pred_dist, [h1.tolist(), f1.tolist()]


def nsm_forward(model, batch):
    """NSM forward with autograd (nsm.py:179-254, forward reasoning only)."""
    local_entity, _qe, facts, q_input, seed_dist, answer_dist = _stage(model, batch)
    B, N = local_entity.shape
    Nt, D = B * N, model.entity_dim
    layer = model.reasoning
    rel_f = model.get_rel_feature_train()
    if model.encode_type:
        h = _type_layer(model.type_layer, facts, rel_f, Nt)
    else:
        h = model.entity_linear(model.entity_embedding(local_entity)).view(Nt, D)
    instructions = _instructions(model.instruction, q_input)
    mask = (local_entity != model.num_entity).float()
    drop = layer.linear_drop_train
    dist = seed_dist
    dist_history = [dist]
    graph = _kernel_graph(model, batch, h.device, D, 1)
    for k in range(model.num_step):
        table = getattr(layer, "rel_linear" + str(k))(rel_f)
        pf = dist.reshape(-1)
        nb = _neighbours(table, None, instructions[:, k:k + 1], dist, facts, graph, Nt).reshape(Nt, D)
        h = F.relu(getattr(layer, "e2e_linear" + str(k))(drop(torch.cat([h, nb], dim=1))))
        m = mask
        if layer.reason_kb:                                                # nsm_gnn.py:98-101
            prior = pf[facts.heads]
            if facts.w is not None:
                prior = prior * facts.w * facts.w
            possible = torch.zeros(Nt, device=h.device).index_add_(0, facts.tails, prior)
            m = mask * (possible > 1e-10).float().view(B, N)
        score = layer.score_func(drop(h)).view(B, N) + (1 - m) * VERY_NEG_NUMBER
        dist = F.softmax(score, dim=1)
        dist_history.append(dist)
    pred_dist = dist_history[-1]
    loss = _loss(model, pred_dist, answer_dist)
    pred = torch.max(pred_dist, dim=1)[1]
    h1, f1 = eval_metric(model, pred_dist.detach(), answer_dist, seed_dist, local_entity)
    model.dist_history = dist_history
    return loss, pred, pred_dist, [h1.tolist(), f1.tolist()]
