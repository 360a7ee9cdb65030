"""TEST INFRASTRUCTURE -- harness that imports the *unmodified* reference (cmavro/GNN-RAG) from
/root/reference/gnn so the oracle restatement can be pinned against it and golden vectors can be
generated (tests/golden/make_golden.py).  /root/reference only exists in the build container; nothing
that runs on the GPU box imports this file.

Shims (reference files untouched):
  * ``BaseInstruction.__init__`` gets a default ``constraint=False`` -- ``LSTMInstruction`` calls
    ``super().__init__(args)`` without it (lstm_encoder.py:13 vs base_encoder.py:10).
  * ``gnn/parsing.py:82`` references an undefined ``create_parser_nutrea`` so ``main.py`` cannot
    start; we bypass the CLI and build the ``args`` dict directly.
  * ``LSTMInstruction`` reads ``<data_folder>/vocab.txt`` (lstm_encoder.py:14); we write a one-line
    file into a temp dir.
"""
import contextlib
import io
import os
import sys
import tempfile
import warnings

REFERENCE_GNN = "/root/reference/gnn"

_state = {}


def available():
    return os.path.isdir(REFERENCE_GNN)


def _import_reference():
    if "mods" in _state:
        return _state["mods"]
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_GNN)
    if REFERENCE_GNN not in sys.path:
        sys.path.insert(0, REFERENCE_GNN)
    warnings.filterwarnings("ignore")
    import modules.question_encoding.base_encoder as be  # noqa: E402

    if not getattr(be.BaseInstruction.__init__, "_gr_shim", False):
        orig = be.BaseInstruction.__init__

        def patched(self, args, constraint=False):
            orig(self, args, constraint)

        patched._gr_shim = True
        be.BaseInstruction.__init__ = patched
    from models.ReaRev.rearev import ReaRev  # noqa: E402
    from models.NSM.nsm import NSM  # noqa: E402
    import evaluate as ref_eval  # noqa: E402
    from modules.kg_reasoning.reasongnn import ReasonGNNLayer  # noqa: E402
    from modules.kg_reasoning.nsm_gnn import NSMLayer  # noqa: E402
    from modules.layer_init import TypeLayer  # noqa: E402

    _state["mods"] = dict(ReaRev=ReaRev, NSM=NSM, evaluate=ref_eval, ReasonGNNLayer=ReasonGNNLayer,
                          NSMLayer=NSMLayer, TypeLayer=TypeLayer)
    return _state["mods"]


def patch_transformers_offline(lm_config):
    """Harness-side stand-in for the hub downloads of bert_encoder.py:31-60,72 (no network here): the tokenizer
    becomes a stub that only knows its pad token id, the encoder is built from ``lm_config`` (a small BertConfig)
    with random weights.  Reference files untouched."""
    import torch
    import transformers

    class _Tok:
        pad_token = "[PAD]"

        def convert_tokens_to_ids(self, tok):
            return 0

    def tok_from_pretrained(name, *a, **k):
        return _Tok()

    def model_from_pretrained(name, *a, **k):
        torch.manual_seed(1234)
        return transformers.AutoModel.from_config(transformers.BertConfig(**lm_config))

    import modules.question_encoding.bert_encoder as be
    be.AutoTokenizer = type("AutoTokenizer", (), {"from_pretrained": staticmethod(tok_from_pretrained)})
    be.AutoModel = type("AutoModel", (), {"from_pretrained": staticmethod(model_from_pretrained)})


def data_folder():
    if "folder" not in _state:
        d = tempfile.mkdtemp(prefix="gr_ref_")
        with open(os.path.join(d, "vocab.txt"), "w") as f:
            f.write("the\n")
        _state["folder"] = d + "/"
    return _state["folder"]


def build_reference_model(args, num_entity, num_relation, num_word, seed=0):
    """Construct the reference nn.Module on CPU with ``torch.manual_seed(seed)`` default init."""
    import torch

    mods = _import_reference()
    args = dict(args)
    args["data_folder"] = data_folder()
    args["use_cuda"] = False
    torch.manual_seed(seed)
    cls = mods[args["model_name"]]
    with contextlib.redirect_stdout(io.StringIO()):
        model = cls(args, num_entity, num_relation, num_word)
    model.eval()
    return model


def reference_forward(model, batch):
    import torch

    with torch.no_grad():
        loss, pred, pred_dist, tp = model(batch[:7])
    return loss, pred, pred_dist


def reference_rank(batch, pred_dist, num_entity, eps):
    """The candidate loop of ``Evaluator.evaluate`` (gnn/evaluate.py:188-209) + ``f1_and_hits``
    (:25-67) run through the reference's own function; returns per-question ordered id/prob lists."""
    mods = _import_reference()
    f1_and_hits = mods["evaluate"].f1_and_hits
    local_entity, query_entities = batch[0], batch[1]
    B, N = local_entity.shape
    ignore_prob = (1 - eps) / N
    id2entity = _Identity()
    out = []
    for b in range(B):
        candidates = local_entity[b].tolist()
        probs = pred_dist[b].tolist()
        seeds = query_entities[b].astype("int64").tolist()
        cand = []
        for c, p, s in zip(candidates, probs, seeds):
            if s == 1.0:
                continue
            if c == num_entity:
                continue
            if p < ignore_prob:
                continue
            cand.append((c, p))
        _, _, _, _, _, _, retrieved, _ = f1_and_hits([], cand, id2entity, None, eps)
        out.append(retrieved)
    return out


class _Identity(dict):
    def __missing__(self, k):
        return k
