"""Generate tests/golden/*.npz by running the UNMODIFIED reference (cmavro/GNN-RAG @ /root/reference)
on seeded synthetic batches.  Run in the build container only:

    python tests/golden/make_golden.py

Each file holds: the args dict (json), the reference state_dict, the batch tuple, and the reference's
outputs (loss, pred, pred_dist, per-iteration dist_history, final node embeddings, the evaluator's
retrieved candidate lists, and one isolated ``reason_layer`` / ``reason_layer_inv`` call).
The vectors pin oracle/kgqa_oracle.py (tests/test_oracle.py) and the CUDA path (tests/test_parity_gpu.py).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gnn_rag_b200 import synthetic as S  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

NUM_ENTITY, NUM_REL, NUM_WORD = 1000, 40, 100


def sharpen(model, e2e=2.0, rel=2.0, score=25.0):
    """Scale a few weight groups so the answer distribution is peaked (default init is ~uniform)."""
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "e2e_linear" in k and k.endswith("weight"):
                p.mul_(e2e)
            if "rel_linear" in k and k.endswith("weight"):
                p.mul_(rel)
            if k.endswith("reasoning.score_func.weight"):
                p.mul_(score)


def add_twins(batch, N):
    """Make local node 5 a structural twin of node 4 in every question (same in/out facts, same
    relative fact order) so the reference produces exact float ties (SURVEY.md §7 hard part 1)."""
    le, qe, kb, qi, sd_, _, ad = batch[:7]
    heads, rels, tails, bids, fids, wl, wrl = kb
    B = le.shape[0]
    h2, r2, t2, b2 = [], [], [], []
    for b in range(B):
        sel = np.where(bids == b)[0]
        a, tw = b * N + 4, b * N + 5
        for f in sel:
            h, r, t = heads[f], rels[f], tails[f]
            if h in (a, tw) or t in (a, tw):
                # drop every fact touching either twin; re-add mirrored pairs below
                continue
            h2.append(h); r2.append(r); t2.append(t); b2.append(b)
        for k, (src, rel) in enumerate([(0, 3), (1, 7), (2, 3)]):
            for node in (a, tw):
                h2.append(b * N + src); r2.append(rel); t2.append(node); b2.append(b)
        for k, (dst, rel) in enumerate([(6, 2), (7, 9)]):
            for node in (a, tw):
                h2.append(node); r2.append(rel); t2.append(b * N + dst); b2.append(b)
        for node in (a, tw):
            h2.append(node); r2.append(NUM_REL - 1); t2.append(node); b2.append(b)
        le[b, 4], le[b, 5] = 777, 778
        ad[b, 4] = ad[b, 5] = 0.0
    heads = np.array(h2, dtype=np.int64); rels = np.array(r2, dtype=np.int64)
    tails = np.array(t2, dtype=np.int64); bids = np.array(b2, dtype=np.int64)
    wl, wrl = S._degree_weights(heads, rels)
    kb = (heads, rels, tails, bids, np.arange(len(heads), dtype=np.int64), wl, wrl)
    return (le, qe, kb, qi, sd_, None, ad) + tuple(batch[7:])


CASES = {
    "rearev_small": dict(model="ReaRev", D=32, kw=dict(num_iter=2, num_ins=2, num_gnn=2),
                         batch=dict(seed=1, B=3, N=50, E=150, n_real="ragged", multi_seed=True)),
    "rearev_norm": dict(model="ReaRev", D=20,
                        kw=dict(num_iter=3, num_ins=3, num_gnn=3, normalized_gnn=True, norm_rel=True),
                        batch=dict(seed=2, B=4, N=64, E=256, n_real="ragged")),
    "rearev_d50_pads": dict(model="ReaRev", D=50, kw=dict(num_iter=3, num_ins=2, num_gnn=3),
                            batch=dict(seed=3, B=4, N=40, E=100, n_real=25, seeds_are_pad=True,
                                       empty_questions=(2,))),
    "rearev_posemb": dict(model="ReaRev", D=24, kw=dict(num_iter=2, num_ins=2, num_gnn=2, pos_emb=True),
                          batch=dict(seed=4, B=2, N=48, E=200, powerlaw=True)),
    "rearev_sharp_ties": dict(model="ReaRev", D=32, kw=dict(num_iter=2, num_ins=2, num_gnn=3),
                              batch=dict(seed=5, B=3, N=60, E=240), sharpen=(4.0, 1.0, 30.0), twins=True),
    "rearev_hub": dict(model="ReaRev", D=16, kw=dict(num_iter=1, num_ins=2, num_gnn=2),
                       batch=dict(seed=6, B=2, N=300, E=6000, powerlaw=True), sharpen=(1.0, 1.0, 0.02)),
    "nsm_small": dict(model="NSM", D=32, kw=dict(num_step=3),
                      batch=dict(seed=7, B=3, N=50, E=150, n_real="ragged", multi_seed=True)),
    "nsm_reason_kb": dict(model="NSM", D=20, kw=dict(num_step=2, reason_kb=True, normalized_gnn=True),
                          batch=dict(seed=8, B=4, N=40, E=120, n_real=30), sharpen=(1.5, 1.5, 50.0)),
    # the configuration of the published checkpoints (gnn/README.md:19: --lm sbert --relation_word_emb True): language-model
    # question encoder + relation-text relation features.  The hub model is replaced by a 1-layer BertConfig of the same
    # width (384) in the harness (no network); token ids: pad = 0 (BERT), words 1..NUM_WORD.
    "rearev_sbert_reltext": dict(model="ReaRev", D=24,
                                 kw=dict(num_iter=2, num_ins=2, num_gnn=2, lm="sbert", relation_word_emb=True,
                                         lm_config=dict(vocab_size=NUM_WORD + 2, hidden_size=384, num_hidden_layers=1,
                                                        num_attention_heads=12, intermediate_size=32,
                                                        max_position_embeddings=32, hidden_dropout_prob=0.0,
                                                        attention_probs_dropout_prob=0.0)),
                                 batch=dict(seed=9, B=3, N=48, E=160, n_real="ragged", multi_seed=True),
                                 sharpen=(2.0, 2.0, 25.0)),
}


def bert_tokens(q_input, num_word):
    """synthetic token ids (pad = num_word) -> BERT convention (pad = 0, words shifted by one)"""
    return np.where(q_input == num_word, 0, q_input + 1).astype(np.int64)


def make_rel_texts(seed, rows, L=5):
    rs = np.random.RandomState(seed)
    t = rs.randint(1, NUM_WORD + 1, size=(rows, L)).astype(np.int64)
    for r in range(rows):
        t[r, rs.randint(2, L + 1):] = 0                      # trailing pads
    return t


def main():
    mods = H._import_reference()
    for name, c in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        args = S.model_args(c["model"], entity_dim=c["D"], word_dim=24, **c["kw"])
        lm = args.get("lm", "lstm") != "lstm"
        if lm:
            H.patch_transformers_offline(args["lm_config"])
        model = H.build_reference_model(args, NUM_ENTITY, NUM_REL, NUM_WORD, seed=0)
        if c.get("sharpen"):
            sharpen(model, *c["sharpen"])
        bkw = dict(c["batch"])
        N = bkw["N"]
        batch = S.make_batch(num_entity=NUM_ENTITY, num_relation=NUM_REL, num_word=NUM_WORD,
                             Q=8, test=True, **bkw)
        if c.get("twins"):
            batch = add_twins(batch, N)
        rel_texts = rel_texts_inv = None
        if lm:
            batch = batch[:3] + (bert_tokens(batch[3], NUM_WORD),) + batch[4:]
            rel_texts, rel_texts_inv = make_rel_texts(31, NUM_REL + 1), make_rel_texts(32, NUM_REL + 1)
            model.encode_rel_texts(rel_texts, rel_texts_inv)                   # gnn/train_model.py:62-64
        loss, pred, pred_dist = H.reference_forward(model, batch)
        retrieved = H.reference_rank(batch, pred_dist, NUM_ENTITY, args["eps"])
        blob = {"args_json": np.array(json.dumps(args))}
        for k, v in model.state_dict().items():
            blob["sd/" + k] = v.detach().numpy()
        le, qe, kb, qi, sdist, _, ad = batch[:7]
        blob.update({"batch/local_entity": le, "batch/query_entities": qe, "batch/q_input": qi,
                     "batch/seed_dist": sdist, "batch/answer_dist": ad,
                     "batch/heads": kb[0], "batch/rels": kb[1], "batch/tails": kb[2],
                     "batch/batch_ids": kb[3], "batch/fact_ids": kb[4],
                     "batch/weight_list": np.array(kb[5], dtype=np.float64),
                     "batch/weight_rel_list": np.array(kb[6], dtype=np.float64)})
        if lm:
            blob["batch/rel_texts"], blob["batch/rel_texts_inv"] = rel_texts, rel_texts_inv
        blob["out/loss"] = loss.numpy()
        blob["out/pred"] = pred.numpy()
        blob["out/pred_dist"] = pred_dist.numpy()
        hist = model.dist_history
        blob["out/dist_history"] = np.stack([h.detach().numpy() for h in hist[1:]])
        blob["out/h_final"] = model.reasoning.local_entity_emb.detach().numpy()
        blob["out/h0"] = model.init_entity_emb.detach().numpy() if hasattr(model, "init_entity_emb") \
            else np.zeros(0, dtype=np.float32)
        ids = [[int(c_) for c_, _ in r] for r in retrieved]
        probs = [[float(p_) for _, p_ in r] for r in retrieved]
        blob["out/cand_len"] = np.array([len(r) for r in ids], dtype=np.int64)
        blob["out/cand_ids"] = np.array(sum(ids, []), dtype=np.int64)
        blob["out/cand_probs"] = np.array(sum(probs, []), dtype=np.float64)
        # one isolated aggregation call through the reference layer (rows 5/6 of SURVEY §8a)
        if c["model"] == "ReaRev":
            layer = model.reasoning
            rs = np.random.RandomState(99)
            B = le.shape[0]
            dist = torch.softmax(torch.from_numpy(rs.randn(B, N).astype(np.float32)), 1)
            ins = torch.from_numpy(rs.randn(B, c["D"]).astype(np.float32))
            with torch.no_grad():
                pe = getattr(layer, "pos_emb0", None) if args.get("pos_emb") else None
                pei = getattr(layer, "pos_emb_inv0", None) if args.get("pos_emb") else None
                nb = layer.reason_layer(dist, ins, layer.rel_linear0, pe)
                nbi = layer.reason_layer_inv(dist, ins, layer.rel_linear0, pei)
            blob["layer/dist"] = dist.numpy()
            blob["layer/ins"] = ins.numpy()
            blob["layer/rel_features"] = layer.rel_features.detach().numpy()
            blob["layer/rel_features_inv"] = layer.rel_features_inv.detach().numpy()
            blob["layer/neighbor_rep"] = nb.numpy()
            blob["layer/neighbor_rep_inv"] = nbi.numpy()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        peak = float(pred_dist.max())
        nties = sum(len(p) - len(set(p)) for p in probs)
        print("%-20s F=%6d loss=%.5f peak_p=%.4f cand=%s ties=%d  %.0f KB" % (
            name, len(kb[0]), float(loss), peak, [len(r) for r in ids], nties,
            os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
