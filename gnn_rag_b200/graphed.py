"""CUDA-graph execution of the whole hot-path step for fixed batch shapes.

One step = CSR batching of the fact list -> model.forward -> candidate ranking.  It launches ~175 kernels, ~130
of them tiny (question encoder, instruction updates, loss), so at WebQSP batch sizes the GPU idles between
launches.  :class:`GraphedStep` captures the step once per input shape ``(B, N, F, Q)`` into a CUDA graph over
static device buffers and replays it: per call it only copies the host batch into the static buffers (H2D from
pinned or pageable memory), replays, and returns views of the static outputs.  Shapes that were not captured yet
are captured on first use; the numerics are those of the eager path (same kernels, same order).
"""
import numpy as np
import torch

from . import batching, ops


class StepOutput:
    __slots__ = ("loss", "pred", "pred_dist", "cand_idx", "cand_count", "cand_total", "db")


class _Captured:
    pass


class GraphedStep:
    def __init__(self, model, num_entity, eps=None):
        self.model = model
        self.num_entity = num_entity
        self.eps = model.eps if eps is None else eps
        self.device = next(model.parameters()).device
        self._cache = {}

    # -- the work that gets captured ------------------------------------------------------------------------
    def _run(self, st):
        m = self.model
        tup = (st.local_entity, st.query_entities, (st.heads, st.rels, st.tails, None, None, None, None),
               st.q_input, st.seed_dist, None, st.answer_dist)
        db = batching.stage_batch(tup, self.device, m.num_relation + 1, False, False)
        loss, pred, pred_dist, _ = m(db)
        cand_idx, cand_count, cand_total = ops.rank_candidates(pred_dist, db.local_entity, db.query_entities,
                                                              self.num_entity, self.eps)
        return db, loss, pred, pred_dist, cand_idx, cand_count, cand_total

    def _capture(self, B, N, F, Q, idx_dtype):
        if self.model.normalized_gnn or self.model.norm_rel:
            raise NotImplementedError("GraphedStep: per-fact weight lists (normalized_gnn / norm_rel) not wired")
        dev = self.device
        st = _Captured()
        st.local_entity = torch.zeros(B, N, dtype=torch.int64, device=dev)
        st.query_entities = torch.zeros(B, N, dtype=torch.float32, device=dev)
        st.seed_dist = torch.zeros(B, N, dtype=torch.float32, device=dev)
        st.answer_dist = torch.zeros(B, N, dtype=torch.float32, device=dev)
        st.q_input = torch.zeros(B, Q, dtype=torch.int64, device=dev)
        st.heads = torch.zeros(F, dtype=idx_dtype, device=dev)
        st.rels = torch.zeros(F, dtype=idx_dtype, device=dev)
        st.tails = torch.zeros(F, dtype=idx_dtype, device=dev)
        return st

    def _fill(self, st, batch):
        le, qe, kb, qi, sd, _, ad = batch[:7]

        def put(dst, src):
            t = src if isinstance(src, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(src))
            if t.dtype != dst.dtype and not t.is_cuda:
                t = t.to(dst.dtype)
            dst.copy_(t, non_blocking=True)
        put(st.local_entity, le); put(st.query_entities, qe); put(st.seed_dist, sd); put(st.answer_dist, ad)
        put(st.q_input, qi); put(st.heads, kb[0]); put(st.rels, kb[1]); put(st.tails, kb[2])

    def __call__(self, batch):
        le, kb, qi = batch[0], batch[2], batch[3]
        B, N = le.shape
        F = int(kb[0].shape[0])
        Q = int(qi.shape[1])
        idx_dtype = torch.int32 if str(kb[0].dtype).endswith("int32") else torch.int64
        # parameter versions are part of the key: the captured graph holds pre-formatted (split-bf16) weights
        key = (B, N, F, Q, idx_dtype, sum(p._version for p in self.model.parameters()))
        ent = self._cache.get(key)
        if ent is None:
            st = self._capture(B, N, F, Q, idx_dtype)
            self._fill(st, batch)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):            # warm-up on a side stream (lazy init, allocator, caches)
                for _ in range(2):
                    self._run(st)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                outs = self._run(st)
            ent = (st, g, outs)
            self._cache[key] = ent
        st, g, outs = ent
        self._fill(st, batch)
        g.replay()
        o = StepOutput()
        o.db, o.loss, o.pred, o.pred_dist, o.cand_idx, o.cand_count, o.cand_total = outs
        self.model.last_batch = o.db
        return o

    def retrieve(self, out):
        """Ordered candidate lists of a :class:`StepOutput` (one D2H), like evaluate.retrieve."""
        from .evaluate import Retrieved
        counts_h = out.cand_count.cpu().numpy()
        maxc = int(counts_h.max()) if counts_h.size else 0
        ei, ef = np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float32)
        if maxc == 0:
            return [Retrieved(ei, ei, ef) for _ in range(out.db.B)], counts_h.size * 4
        idx = out.cand_idx[:, :maxc].long()
        probs = torch.gather(out.pred_dist, 1, idx)
        ents = torch.gather(out.db.local_entity, 1, idx)
        idx_h, probs_h, ents_h = idx.cpu().numpy(), probs.cpu().numpy(), ents.cpu().numpy()
        res = [Retrieved(idx_h[b, :c], ents_h[b, :c], probs_h[b, :c]) for b, c in enumerate(counts_h.tolist())]
        return res, counts_h.size * 4 + idx_h.size * 8 + probs_h.size * 4 + ents_h.size * 8
