"""TEST INFRASTRUCTURE -- CPU oracle for the GNN retrieval hot path of cmavro/GNN-RAG.

A functional (state_dict in, tensors out) restatement of the reference's ReaRev / NSM forward,
written against torch-CPU fp32 because the path is floating point.  It deliberately keeps the
reference's *operation order and redundancy* (per-fact ``index_select`` -> ``Linear`` over F rows ->
COO ``sparse.mm``) so that (i) numerics track the reference bit-for-bit where torch is deterministic
and (ii) timing it is an honest stand-in for the reference's CPU path (bench.py cpu_baseline,
kind="port").  It is NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.

Parity pin: tests/golden/*.npz were produced by running the unmodified reference
(oracle/ref_harness.py + tests/golden/make_golden.py) in the build container; tests/test_oracle.py
checks this file against them (pred_dist to 1e-6 relative, candidate lists exactly).

Every function cites the reference lines it follows (paths under /root/reference/).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

VERY_NEG_NUMBER = -100000000000  # reasongnn.py:9, nsm_gnn.py:12, base_encoder.py:6
VERY_SMALL_NUMBER = 1e-10        # nsm_gnn.py:11


# --------------------------------------------------------------------------------------------
# sparse structure  (gnn/modules/kg_reasoning/base_gnn.py:19-54)
# --------------------------------------------------------------------------------------------
class FactMats:
    """The COO operators ``build_matrix`` creates, restricted to the four the forward reads."""

    def __init__(self, kb_adj_mat, B, N, normalized_gnn):
        heads, rels, tails, bids, fact_ids, weight_list, _ = kb_adj_mat
        Fn = len(fact_ids)
        self.num_fact = Fn
        self.batch_rels = torch.from_numpy(np.asarray(rels, dtype=np.int64))
        self.batch_ids = torch.from_numpy(np.asarray(bids, dtype=np.int64))
        h = torch.from_numpy(np.asarray(heads, dtype=np.int64))
        t = torch.from_numpy(np.asarray(tails, dtype=np.int64))
        f = torch.from_numpy(np.asarray(fact_ids, dtype=np.int64))
        if normalized_gnn:                                   # base_gnn.py:38-41
            vals = torch.tensor(weight_list, dtype=torch.float32)
        else:
            vals = torch.ones(Fn, dtype=torch.float32)
        Nt = B * N
        sp = torch.sparse_coo_tensor                          # base_gnn.py:53-54 (uncoalesced COO)
        self.fact2head = sp(torch.stack([h, f]), vals, (Nt, Fn))
        self.head2fact = sp(torch.stack([f, h]), vals, (Fn, Nt))
        self.fact2tail = sp(torch.stack([t, f]), vals, (Nt, Fn))
        self.tail2fact = sp(torch.stack([f, t]), vals, (Fn, Nt))


# --------------------------------------------------------------------------------------------
# TypeLayer  (gnn/modules/layer_init.py:25-62)
# --------------------------------------------------------------------------------------------
def type_layer(sd, prefix, kb_adj_mat, rel_features, B, N, norm_rel):
    heads, rels, tails, bids, fact_ids, _, weight_rel_list = kb_adj_mat
    Fn = len(fact_ids)
    h = torch.from_numpy(np.asarray(heads, dtype=np.int64))
    t = torch.from_numpy(np.asarray(tails, dtype=np.int64))
    f = torch.from_numpy(np.asarray(fact_ids, dtype=np.int64))
    r = torch.from_numpy(np.asarray(rels, dtype=np.int64))
    if norm_rel:                                              # layer_init.py:39-42
        val_one = torch.tensor(weight_rel_list, dtype=torch.float32)
    else:
        val_one = torch.ones(Fn, dtype=torch.float32)
    fact_rel = torch.index_select(rel_features, 0, r)         # :46
    fact_val = F.linear(fact_rel, sd[prefix + "kb_self_linear.weight"],
                        sd[prefix + "kb_self_linear.bias"])   # :48
    Nt = B * N
    f2t = torch.sparse_coo_tensor(torch.stack([t, f]), val_one, (Nt, Fn))   # :52
    f2h = torch.sparse_coo_tensor(torch.stack([h, f]), val_one, (Nt, Fn))   # :53
    out = F.relu(torch.sparse.mm(f2t, fact_val) + torch.sparse.mm(f2h, fact_val))  # :57
    return out.view(B, N, -1)


# --------------------------------------------------------------------------------------------
# question encoder + instruction attention
# (gnn/modules/question_encoding/lstm_encoder.py:32-45, base_encoder.py:73-114)
# --------------------------------------------------------------------------------------------
class LstmQuestion:
    def __init__(self, sd, q_input, num_word, D):
        p = "instruction."
        emb = F.embedding(q_input, sd["word_embedding.weight"])             # lstm_encoder.py:34
        lstm = torch.nn.LSTM(emb.shape[-1], D, batch_first=True)
        with torch.no_grad():
            lstm.weight_ih_l0.copy_(sd[p + "node_encoder.weight_ih_l0"])
            lstm.weight_hh_l0.copy_(sd[p + "node_encoder.weight_hh_l0"])
            lstm.bias_ih_l0.copy_(sd[p + "node_encoder.bias_ih_l0"])
            lstm.bias_hh_l0.copy_(sd[p + "node_encoder.bias_hh_l0"])
        lstm.eval()
        Bq = q_input.shape[0]
        h0 = torch.zeros(1, Bq, D)
        with torch.no_grad():
            hidden, (h_n, _) = lstm(emb, (h0, h0.clone()))                  # :35-37
        self.query_hidden_emb = hidden                                       # :42
        self.query_node_emb = h_n.squeeze(0).unsqueeze(1)                    # :41
        self.query_mask = (q_input != num_word).float()                      # :43
        self.sd = sd
        self.D = D

    def get_instruction(self, relational_ins, step):                         # base_encoder.py:82-101
        sd, p = self.sd, "instruction."
        ri = relational_ins.unsqueeze(1)
        q_i = F.linear(self.query_node_emb, sd[p + "question_linear%d.weight" % step],
                       sd[p + "question_linear%d.bias" % step])              # :93
        cq = F.linear(torch.cat((ri, q_i, q_i - ri, q_i * ri), dim=-1),
                      sd[p + "cq_linear.weight"], sd[p + "cq_linear.bias"])  # :94
        ca = F.linear(cq * self.query_hidden_emb, sd[p + "ca_linear.weight"],
                      sd[p + "ca_linear.bias"])                              # :96
        attn = F.softmax(ca + (1 - self.query_mask.unsqueeze(2)) * VERY_NEG_NUMBER, dim=1)  # :99
        return torch.sum(attn * self.query_hidden_emb, dim=1)                # :101

    def run(self, num_ins):                                                  # base_encoder.py:105-114
        ri = torch.zeros(self.query_hidden_emb.shape[0], self.D)             # :77
        out = []
        for i in range(num_ins):
            ri = self.get_instruction(ri, i)
            out.append(ri)
        return out


# --------------------------------------------------------------------------------------------
# ReaRev reasoning layer  (gnn/modules/kg_reasoning/reasongnn.py:61-174)
# --------------------------------------------------------------------------------------------
def reason_layer(mats, curr_dist, instruction, rel_features, W, b, inverse, pos_emb=None):
    """reasongnn.py:61-89 (inverse=False) / :91-116 (inverse=True)."""
    fact_rel = torch.index_select(rel_features, 0, mats.batch_rels)          # :71 / :97
    fact_query = torch.index_select(instruction, 0, mats.batch_ids)          # :73 / :99
    lin = F.linear(fact_rel, W, b)
    if pos_emb is not None:
        lin = lin + F.embedding(mats.batch_rels, pos_emb)                    # :75-77
    fact_val = F.relu(lin * fact_query)                                      # :79 / :105
    src = mats.tail2fact if inverse else mats.head2fact
    dst = mats.fact2head if inverse else mats.fact2tail
    fact_prior = torch.sparse.mm(src, curr_dist.reshape(-1, 1))              # :80 / :106
    fact_val = fact_val * fact_prior                                         # :82 / :109
    return torch.sparse.mm(dst, fact_val)                                    # :84 / :111


def rearev_gnn_step(sd, mats, h, curr_dist, relational_ins, rel_f, rel_f_inv, mask, step,
                    pos_emb=False):
    """ReasonGNNLayer.forward, reasongnn.py:134-174.  h: [B,N,D]; returns (dist, h_new, score)."""
    B, N, D = h.shape
    p = "reasoning."
    W, b = sd[p + "rel_linear%d.weight" % step], sd[p + "rel_linear%d.bias" % step]
    pe = sd[p + "pos_emb%d.weight" % step] if pos_emb else None
    pei = sd[p + "pos_emb_inv%d.weight" % step] if pos_emb else None
    reps = []
    for j in range(relational_ins.shape[1]):                                 # :150-156
        reps.append(reason_layer(mats, curr_dist, relational_ins[:, j, :], rel_f, W, b, False,
                                 pe).view(B, N, D))
        reps.append(reason_layer(mats, curr_dist, relational_ins[:, j, :], rel_f_inv, W, b, True,
                                 pei).view(B, N, D))
    x = torch.cat([h] + reps, dim=2)                                         # :158-161
    h_new = F.relu(F.linear(x, sd[p + "e2e_linear%d.weight" % step],
                            sd[p + "e2e_linear%d.bias" % step]))             # :163
    score = F.linear(h_new, sd[p + "score_func.weight"], sd[p + "score_func.bias"]).squeeze(2)  # :165
    score = score + (1 - mask) * VERY_NEG_NUMBER                             # :168
    return F.softmax(score, dim=1), h_new, score                             # :169


# --------------------------------------------------------------------------------------------
# instruction update  (gnn/modules/query_update.py:6-44)
# --------------------------------------------------------------------------------------------
def query_reform(sd, prefix, q_node, ent_emb, seed_info):
    seed_retrieve = torch.bmm(seed_info.unsqueeze(1), ent_emb).squeeze(1)    # :40
    x, y = q_node, seed_retrieve
    cat = torch.cat([x, y, x - y], dim=-1)
    r_ = F.linear(cat, sd[prefix + "fusion.r.weight"])                       # :13
    g_ = torch.sigmoid(F.linear(cat, sd[prefix + "fusion.g.weight"]))        # :14
    return g_ * r_ + (1 - g_) * x                                            # :15


# --------------------------------------------------------------------------------------------
# loss  (gnn/models/base_model.py:193-215, rearev.py:156-160)
# --------------------------------------------------------------------------------------------
def kl_loss(pred_dist, answer_dist):
    answer_len = torch.sum(answer_dist, dim=1, keepdim=True)
    case_valid = (answer_len > 0).float()                                    # rearev.py:229-230
    answer_len = answer_len.clone()
    answer_len[answer_len == 0] = 1.0                                        # base_model.py:195
    answer_prob = answer_dist / answer_len
    log_prob = torch.log(pred_dist + 1e-8)                                   # :197
    tp = F.kl_div(log_prob, answer_prob, reduction="none")                   # :198
    return torch.sum(tp * case_valid) / pred_dist.shape[0]                   # rearev.py:158-159


def _unpack(batch):
    local_entity, query_entities, kb_adj_mat, q_input, seed_dist, _, answer_dist = batch[:7]
    le = torch.from_numpy(local_entity).long()
    qe = torch.from_numpy(query_entities).float()
    ad = torch.from_numpy(answer_dist).float()
    sdist = torch.from_numpy(seed_dist).float()
    qi = torch.from_numpy(q_input).long()
    return le, qe, kb_adj_mat, qi, sdist, ad


# --------------------------------------------------------------------------------------------
# ReaRev.forward  (gnn/models/ReaRev/rearev.py:163-243)
# --------------------------------------------------------------------------------------------
def rearev_forward(sd, args, num_entity, num_word, batch, return_trace=False):
    """Returns (loss, pred, pred_dist[, trace]).  ``sd`` = reference state_dict (fp32 CPU tensors)."""
    D, T, K, I = args["entity_dim"], args["num_iter"], args["num_gnn"], args["num_ins"]
    le, qe, kb, qi, seed_dist, answer_dist = _unpack(batch)
    B, N = le.shape
    # init_reason, rearev.py:132-153
    rel_f = F.linear(sd["relation_embedding.weight"], sd["relation_linear.weight"],
                     sd["relation_linear.bias"])                             # :96-100
    rel_f_inv = F.linear(sd["relation_embedding_inv.weight"], sd["relation_linear.weight"],
                         sd["relation_linear.bias"])
    h = type_layer(sd, "type_layer.", kb, rel_f, B, N, args["norm_rel"])     # :81-84
    mask = (le != num_entity).float()                                        # reasongnn.py:48
    mats = FactMats(kb, B, N, args["normalized_gnn"])                        # reasongnn.py:57
    # the question is encoded a second time and *these* instructions are the ones used (:192-196)
    q = LstmQuestion(sd, qi, num_word, D)
    instructions = q.run(I)
    trace = dict(h0=h.clone(), rel_f=rel_f, rel_f_inv=rel_f_inv,
                 instructions=[x.clone() for x in instructions], dists=[], neighbor_reps=None)
    dist = seed_dist
    for t in range(T):                                                       # :206
        relation_ins = torch.stack(instructions, dim=1)                      # :207
        dist = seed_dist                                                     # :208 (reset to seed)
        for j in range(K):                                                   # :209-210
            dist, h, _ = rearev_gnn_step(sd, mats, h, dist, relation_ins, rel_f, rel_f_inv, mask,
                                         j, args.get("pos_emb", False))
            trace["dists"].append(dist.clone())
        for j in range(I):                                                   # :217-221
            instructions[j] = query_reform(sd, "reform%d." % j, instructions[j], h, qe)
    loss = kl_loss(dist, answer_dist)                                        # :233
    pred = torch.max(dist, dim=1)[1]                                         # :237
    trace["h_final"] = h
    if return_trace:
        return loss, pred, dist, trace
    return loss, pred, dist


# --------------------------------------------------------------------------------------------
# NSM  (gnn/modules/kg_reasoning/nsm_gnn.py:54-112, gnn/models/NSM/nsm.py:179-254)
# --------------------------------------------------------------------------------------------
def nsm_gnn_step(sd, mats, h, curr_dist, instruction, rel_f, mask, step, reason_kb):
    B, N, D = h.shape
    p = "reasoning."
    W, b = sd[p + "rel_linear%d.weight" % step], sd[p + "rel_linear%d.bias" % step]
    fact_rel = torch.index_select(rel_f, 0, mats.batch_rels)                 # nsm_gnn.py:93
    fact_query = torch.index_select(instruction, 0, mats.batch_ids)          # :96
    fact_val = F.relu(F.linear(fact_rel, W, b) * fact_query)                 # :97
    fact_prior = torch.sparse.mm(mats.head2fact, curr_dist.reshape(-1, 1))   # :98
    possible_tail = torch.sparse.mm(mats.fact2tail, fact_prior)              # :101
    possible_tail = (possible_tail > VERY_SMALL_NUMBER).float().view(B, N)   # :103
    fact_val = fact_val * fact_prior                                         # :105
    nb = torch.sparse.mm(mats.fact2tail, fact_val).view(B, N, D)             # :107-110
    x = torch.cat((h, nb), dim=2)                                            # nsm_gnn.py:62
    h_new = F.relu(F.linear(x, sd[p + "e2e_linear%d.weight" % step],
                            sd[p + "e2e_linear%d.bias" % step]))             # :63-65
    score = F.linear(h_new, sd[p + "score_func.weight"], sd[p + "score_func.bias"]).squeeze(2)  # :67
    answer_mask = mask * possible_tail if reason_kb else mask                # :68-71
    score = score + (1 - answer_mask) * VERY_NEG_NUMBER                      # :73
    return F.softmax(score, dim=1), h_new, possible_tail                     # :74


def nsm_forward(sd, args, num_entity, num_word, batch, return_trace=False):
    D, S = args["entity_dim"], args["num_step"]
    le, qe, kb, qi, seed_dist, answer_dist = _unpack(batch)
    B, N = le.shape
    q = LstmQuestion(sd, qi, num_word, D)
    instruction_list = q.run(S)                                              # nsm.py:117
    rel_f = F.linear(sd["relation_embedding.weight"], sd["relation_linear1.weight"],
                     sd["relation_linear1.bias"])                            # nsm.py:98-100
    h = type_layer(sd, "type_layer.", kb, rel_f, B, N, args["norm_rel"])     # nsm.py:121
    mask = (le != num_entity).float()
    mats = FactMats(kb, B, N, args["normalized_gnn"])
    dist = seed_dist
    trace = dict(h0=h.clone(), dists=[], possible_tail=[])
    for i in range(S):                                                       # nsm.py:219-222
        dist, h, pt = nsm_gnn_step(sd, mats, h, dist, instruction_list[i], rel_f, mask, i,
                                   args.get("reason_kb", False))
        trace["dists"].append(dist.clone())
        trace["possible_tail"].append(pt)
    loss = kl_loss(dist, answer_dist)                                        # nsm.py:244
    pred = torch.max(dist, dim=1)[1]
    trace["h_final"] = h
    if return_trace:
        return loss, pred, dist, trace
    return loss, pred, dist


def forward(sd, args, num_entity, num_word, batch, **kw):
    if args["model_name"] == "NSM":
        return nsm_forward(sd, args, num_entity, num_word, batch, **kw)
    return rearev_forward(sd, args, num_entity, num_word, batch, **kw)


# --------------------------------------------------------------------------------------------
# candidate ranking  (gnn/evaluate.py:156, 188-209 and f1_and_hits :25-67)  -- pure Python loops
# --------------------------------------------------------------------------------------------
def rank_candidates(local_entity, query_entities, pred_dist, num_entity, eps):
    """Per question: ordered list of (local_index, entity_id, prob) the evaluator retrieves.

    Drop seeds (``s == 1.0``), pads (``c == pad``) and ``p < (1-eps)/N``; *stable* sort by p
    descending (python ``sorted(..., reverse=True)`` keeps equal keys in original order,
    evaluate.py:34); take the prefix up to and including the item whose cumulative float64 sum
    exceeds eps (:41-50).
    """
    B, N = local_entity.shape
    ignore_prob = (1 - eps) / N                                              # evaluate.py:156
    probs_all = np.asarray(pred_dist, dtype=np.float32)
    out = []
    for b in range(B):
        cand = []
        probs = probs_all[b].tolist()                                        # fp32 -> python float
        cands = local_entity[b].tolist()
        seeds = np.asarray(query_entities[b]).astype(np.int64).tolist()      # evaluate.py:173
        for n in range(N):
            if seeds[n] == 1.0:
                continue
            if cands[n] == num_entity:
                continue
            if probs[n] < ignore_prob:
                continue
            cand.append((n, cands[n], probs[n]))
        cand = sorted(cand, key=lambda x: x[2], reverse=True)
        tp = 0.0
        ret = []
        for n, c, p in cand:
            ret.append((n, c, p))
            tp += p
            if tp > eps:
                break
        out.append(ret)
    return out


def f1_and_hits(answers, retrieved_ids, best_ans):
    """Metric tail of evaluate.py:51-67 on an already-cut candidate list."""
    correct = sum(1 for c in retrieved_ids if c in answers)
    if len(answers) == 0:
        return (1.0, 1.0, 1.0, 1.0) if len(retrieved_ids) == 0 else (0.0, 1.0, 0.0, 1.0)
    hits = float(best_ans in answers)
    if len(retrieved_ids) == 0:
        return 1.0, 0.0, 0.0, hits
    p, r = correct / len(retrieved_ids), correct / len(answers)
    f1 = 2.0 / (1.0 / p + 1.0 / r) if p != 0 and r != 0 else 0.0
    return p, r, f1, hits


# --------------------------------------------------------------------------------------------
# shortest-path node sets  (llm/src/utils/graph_utils.py:10-21, 49-75) -- SURVEY.md §8f row 1
# --------------------------------------------------------------------------------------------
def shortest_path_nodes(heads, tails, n_nodes, sources, targets):
    """Set of nodes lying on *any* shortest path between any source and any target in the
    undirected graph (``nx.Graph`` + ``nx.all_shortest_paths``).  Plain BFS, small inputs only.
    Returns (sorted node list, dict target -> hop distance from the nearest... per (s,t) pair).
    """
    adj = [[] for _ in range(n_nodes)]
    for h, t in zip(heads, tails):
        if h != t:
            adj[h].append(t)
            adj[t].append(h)

    def bfs(src):
        dist = [-1] * n_nodes
        dist[src] = 0
        q = [src]
        for u in q:
            for v in adj[u]:
                if dist[v] < 0:
                    dist[v] = dist[u] + 1
                    q.append(v)
        return dist

    nodes = set()
    pair_dist = {}
    tdist = {t: bfs(t) for t in set(targets)}
    for s in sources:
        ds = bfs(s)
        for t in targets:
            if ds[t] < 0:
                continue
            pair_dist[(s, t)] = ds[t]
            dt = tdist[t]
            for v in range(n_nodes):
                if ds[v] >= 0 and dt[v] >= 0 and ds[v] + dt[v] == ds[t]:
                    nodes.add(v)
    return sorted(nodes), pair_dist
