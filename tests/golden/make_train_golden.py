"""Generate tests/golden/train/*.npz: loss, train-time metrics and parameter gradients of
``model(batch, training=True)`` + ``loss.backward()`` computed by the UNMODIFIED reference (cmavro/GNN-RAG @
/root/reference), on the weights and batches of the forward goldens (tests/golden/*.npz).  The reference model is in
eval() mode so that dropout is the identity and the numbers are deterministic (gnn/train_model.py:209-233 runs the
same call in train() mode).  Run in the build container only:

    python tests/golden/make_train_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_io import Golden  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "train")
CASES = ["rearev_small", "rearev_norm", "rearev_posemb", "rearev_sharp_ties", "nsm_small", "nsm_reason_kb",
         "rearev_sbert_reltext"]
HIT_CASES = {"rearev_sharp_ties", "nsm_reason_kb", "rearev_small"}   # answers moved onto the top-1 node: h1 = 1, f1 > 0


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in CASES:
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        g = Golden(name)
        if g.args.get("lm", "lstm") != "lstm":
            H._import_reference()
            H.patch_transformers_offline(g.args["lm_config"])
        model = H.build_reference_model(g.args, g.num_entity, g.num_relation, g.num_word, seed=0)
        model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in g.sd.items()}, strict=True)
        model.eval()
        if g.rel_texts is not None:
            model.encode_rel_texts(g.rel_texts, g.rel_texts_inv)
        for k, p in model.named_parameters():
            if "node_encoder" not in k:                  # the language model stays frozen (lm_frozen = 1)
                p.requires_grad_(True)
        batch = list(g.batch[:7])
        if name in HIT_CASES:      # random-init models never hit: make the forward golden's top-1 node (+ node 7) the answers
            ad = np.zeros_like(batch[6])
            top = g.out["pred_dist"].argmax(1)
            for b, t in enumerate(top):
                ad[b, t] = 1.0
                ad[b, 7] = 1.0
            batch[6] = ad
        loss, pred, pred_dist, tp_list = model(tuple(batch), training=True)
        loss.backward()
        blob = {"answer_dist": batch[6], "loss": loss.detach().numpy(), "pred_dist": pred_dist.detach().numpy(),
                "h1": np.array(tp_list[0], dtype=np.float32), "f1": np.array(tp_list[1], dtype=np.float32)}
        n = 0
        for k, p in model.named_parameters():
            if p.grad is not None:
                blob["grad/" + k] = p.grad.numpy()
                n += 1
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print("%-20s loss=%.6f h1=%s f1=%s grads=%d  %.0f KB" % (name, float(loss), tp_list[0], tp_list[1], n,
                                                                os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
