"""Load tests/golden/*.npz (made by tests/golden/make_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix="", include_lm=False):
    """Golden case names.  ``include_lm``: also the language-model-encoder cases (``*_sbert_*``), which the CPU
    oracle port does not restate (the HuggingFace encoder is the input of the path)."""
    out = sorted(os.path.splitext(os.path.basename(p))[0]
                 for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in out if n.startswith(prefix) and (include_lm or "sbert" not in n)]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.args = json.loads(str(z["args_json"]))
        self.sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
        b = {k[6:]: z[k] for k in z.files if k.startswith("batch/")}
        kb = (b["heads"], b["rels"], b["tails"], b["batch_ids"], b["fact_ids"],
              b["weight_list"].tolist(), b["weight_rel_list"].tolist())
        self.batch = (b["local_entity"], b["query_entities"], kb, b["q_input"], b["seed_dist"], None,
                      b["answer_dist"])
        self.rel_texts = b.get("rel_texts")
        self.rel_texts_inv = b.get("rel_texts_inv")
        self.out = {k[4:]: z[k] for k in z.files if k.startswith("out/")}
        self.layer = {k[6:]: z[k] for k in z.files if k.startswith("layer/")}
        self.num_entity, self.num_relation, self.num_word = 1000, 40, 100

    def cand_lists(self):
        lens = self.out["cand_len"].tolist()
        ids, probs, o = [], [], 0
        for n in lens:
            ids.append(self.out["cand_ids"][o:o + n].tolist())
            probs.append(self.out["cand_probs"][o:o + n].tolist())
            o += n
        return ids, probs
