"""CUDA-graph execution of the whole hot-path step for fixed batch shapes.

One step = CSR batching of the fact list -> model.forward -> candidate ranking.  It launches ~70 kernels, most of
them small, so at WebQSP batch sizes the GPU idles between launches.  :class:`GraphedStep` captures the step into a
CUDA graph over static device buffers and replays it: per call it only copies the host batch into the static buffers
(H2D from pinned or pageable memory), replays, and returns views of the static outputs.  The numerics are those of
the eager path (same kernels, same order).

Shapes.  A graph is fixed in ``(B, N, Q)`` and in the CAPACITY of its fact buffers.  ``get_batch`` returns a different
fact count F for almost every batch (gnn/dataset_load.py:473-527), so capacities are bucketed (8 buckets per octave,
<= 12.5 % padding): the batch's facts occupy the front of the buffers, a device-side counter tells the CSR build how
many slots are live (``gr_csr_build(..., nfacts)``) and everything downstream sees live facts only, through the row
pointers.  Captured graphs are kept in an LRU cache (``max_graphs``); the graph, its static buffers, landing buffers
and pinned host buffers of an evicted entry are released.

Serving loop: :meth:`GraphedStep.submit` / :meth:`GraphedStep.collect` pipeline two batches -- the H2D copy of
batch i+1 (copy stream, into a landing buffer set) and the D2H read of batch i's results overlap the graph of
batch i, so the end-to-end rate is bounded by the device time of the step, not by device + PCIe time.
"""
import collections

import numpy as np
import torch

from . import batching, ops


class StepOutput:
    __slots__ = ("loss", "pred", "pred_dist", "cand_idx", "cand_count", "cand_total", "db")


class _Captured:
    pass


class Ticket:
    """One in-flight step of the submit/collect pipeline."""
    __slots__ = ("slot", "done", "ent", "local_entity_host", "B", "N")


def fact_capacity(F):
    """Bucketed capacity for a batch of F facts: next multiple of 2^(floor(log2 F) - 3), at least 1024."""
    F = max(int(F), 1)
    g = max(1 << max(F.bit_length() - 4, 0), 1024)
    return (F + g - 1) // g * g


class GraphedStep:
    def __init__(self, model, num_entity, eps=None, max_graphs=8):
        self.model = model
        self.num_entity = num_entity
        self.eps = model.eps if eps is None else eps
        self.device = next(model.parameters()).device
        self.max_graphs = max_graphs
        self._cache = collections.OrderedDict()
        self._copy_stream = None      # H2D stream
        self._d2h_stream = None       # separate: a D2H waiting for graph i must not block the H2D of batch i+1
        self._slot = 0
        self._weights = bool(model.normalized_gnn), bool(model.norm_rel)

    # -- the work that gets captured ------------------------------------------------------------------------
    def _run(self, st):
        m = self.model
        tup = (st.local_entity, st.query_entities,
               (st.heads, st.rels, st.tails, None, None, st.weight_list, st.weight_rel_list),
               st.q_input, st.seed_dist, None, st.answer_dist)
        db = batching.stage_batch(tup, self.device, m.num_relation + 1, m.normalized_gnn, m.norm_rel,
                                  nfacts=st.nfacts)
        loss, pred, pred_dist, _ = m(db)
        cand_idx, cand_count, cand_total = ops.rank_candidates(pred_dist, db.local_entity, db.query_entities,
                                                              self.num_entity, self.eps)
        return db, loss, pred, pred_dist, cand_idx, cand_count, cand_total

    def _static_inputs(self, B, N, cap, Q, idx_dtype):
        dev = self.device
        st = _Captured()
        st.local_entity = torch.zeros(B, N, dtype=torch.int64, device=dev)
        st.query_entities = torch.zeros(B, N, dtype=torch.float32, device=dev)
        st.seed_dist = torch.zeros(B, N, dtype=torch.float32, device=dev)
        st.answer_dist = torch.zeros(B, N, dtype=torch.float32, device=dev)
        st.q_input = torch.zeros(B, Q, dtype=torch.int64, device=dev)
        st.heads = torch.zeros(cap, dtype=idx_dtype, device=dev)
        st.rels = torch.zeros(cap, dtype=idx_dtype, device=dev)
        st.tails = torch.zeros(cap, dtype=idx_dtype, device=dev)
        st.nfacts = torch.zeros(1, dtype=torch.int32, device=dev)
        st.weight_list = torch.ones(cap, dtype=torch.float32, device=dev) if self._weights[0] else None
        st.weight_rel_list = torch.ones(cap, dtype=torch.float32, device=dev) if self._weights[1] else None
        return st

    def _names(self):
        names = ["local_entity", "query_entities", "seed_dist", "answer_dist", "q_input", "heads", "rels", "tails",
                 "nfacts"]
        if self._weights[0]:
            names.append("weight_list")
        if self._weights[1]:
            names.append("weight_rel_list")
        return names

    def _fill(self, st, batch):
        """Host batch -> the buffers of ``st`` (facts to the front of the capacity, live count to ``nfacts``)."""
        le, qe, kb, qi, sd, _, ad = batch[:7]
        F = int(kb[0].shape[0])

        def host(src, dtype):
            t = src if isinstance(src, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(src)))
            if t.dtype != dtype and not t.is_cuda:
                t = t.to(dtype)
            return t

        def put(dst, src):
            dst.copy_(host(src, dst.dtype), non_blocking=True)

        def put_front(dst, src):
            dst[:F].copy_(host(src, dst.dtype), non_blocking=True)
        put(st.local_entity, le); put(st.query_entities, qe); put(st.seed_dist, sd); put(st.answer_dist, ad)
        put(st.q_input, qi)
        put_front(st.heads, kb[0]); put_front(st.rels, kb[1]); put_front(st.tails, kb[2])
        if st.weight_list is not None:
            if kb[5] is None:
                raise ValueError("normalized_gnn needs kb_adj_mat's weight_list")
            put_front(st.weight_list, np.asarray(kb[5], dtype=np.float32) if not isinstance(kb[5], torch.Tensor) else kb[5])
        if st.weight_rel_list is not None:
            if kb[6] is None:
                raise ValueError("norm_rel needs kb_adj_mat's weight_rel_list")
            put_front(st.weight_rel_list,
                      np.asarray(kb[6], dtype=np.float32) if not isinstance(kb[6], torch.Tensor) else kb[6])
        st.nfacts.copy_(torch.tensor([F], dtype=torch.int32), non_blocking=True)
        idx_b = st.heads.element_size()
        self._h2d_bytes = (st.local_entity.numel() * 8 + st.q_input.numel() * 8 + 3 * st.seed_dist.numel() * 4
                           + 3 * F * idx_b + 4 + 4 * F * (int(st.weight_list is not None) +
                                                          int(st.weight_rel_list is not None)))

    @staticmethod
    def _needs_staging(batch):
        """True when the batch lives in pageable host memory (numpy arrays / unpinned CPU tensors)."""
        x = batch[2][0]
        if isinstance(x, torch.Tensor):
            return (not x.is_cuda) and (not x.is_pinned())
        return True

    def _stage_host(self, stage, batch):
        """Cast + copy a pageable ``get_batch`` tuple into the pinned staging set (host memcpy); returns a tuple over the
        staged tensors that ``_fill`` can DMA asynchronously."""
        le, qe, kb, qi, sd, _, ad = batch[:7]
        F = int(kb[0].shape[0])

        def put(dst, src, n=None):
            d = dst.numpy() if n is None else dst.numpy()[:n]
            np.copyto(d, src.numpy() if isinstance(src, torch.Tensor) else np.asarray(src), casting="unsafe")
            return dst if n is None else dst[:n]
        wl = put(stage.weight_list, np.asarray(kb[5], dtype=np.float32), F) if stage.weight_list is not None else None
        wr = put(stage.weight_rel_list, np.asarray(kb[6], dtype=np.float32), F) if stage.weight_rel_list is not None \
            else None
        kb2 = (put(stage.heads, kb[0], F), put(stage.rels, kb[1], F), put(stage.tails, kb[2], F), None, None, wl, wr)
        return (put(stage.local_entity, le), put(stage.query_entities, qe), kb2, put(stage.q_input, qi),
                put(stage.seed_dist, sd), None, put(stage.answer_dist, ad))

    def _entry(self, batch):
        le, kb, qi = batch[0], batch[2], batch[3]
        B, N = le.shape
        cap = fact_capacity(int(kb[0].shape[0]))
        Q = int(qi.shape[1])
        idx_dtype = torch.int32 if str(kb[0].dtype).endswith("int32") else torch.int64
        # parameter versions are part of the key: the captured graph holds pre-formatted (split-bf16) weights
        key = (B, N, cap, Q, idx_dtype, sum(p._version for p in self.model.parameters()))
        ent = self._cache.get(key)
        if ent is not None:
            self._cache.move_to_end(key)
            return ent
        while len(self._cache) >= self.max_graphs:          # LRU eviction: graph, static + landing + pinned buffers
            _k, old = self._cache.popitem(last=False)
            torch.cuda.synchronize()
            del old
        st = self._static_inputs(B, N, cap, Q, idx_dtype)
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
        ent = _Captured()
        ent.st, ent.g, ent.outs = st, g, outs
        ent.pipe = None
        ent.weight_ws = ops.live_weight_workspaces()        # the graph reads these pre-formatted weights: keep them alive
        self._cache[key] = ent
        return ent

    @staticmethod
    def _check(db):
        """ids outside the batch are clamped by the CSR build and flagged (a malformed / mis-sharded fact list)."""
        db.graph.check_status()

    def __call__(self, batch, check=False):
        ent = self._entry(batch)
        self._fill(ent.st, batch)
        ent.g.replay()
        o = StepOutput()
        o.db, o.loss, o.pred, o.pred_dist, o.cand_idx, o.cand_count, o.cand_total = ent.outs
        o.db.h2d_bytes = self._h2d_bytes
        self.model.last_batch = o.db
        if check:
            self._check(o.db)
        return o

    # -- two-deep pipeline: H2D of batch i+1 and D2H of batch i overlap the graph of batch i -----------------
    def _pipe(self, ent):
        if ent.pipe is None:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
                self._d2h_stream = torch.cuda.Stream()
            db, loss, pred, pred_dist, cand_idx, cand_count, _ = ent.outs
            pipe = _Captured()
            pipe.land, pipe.out_dev, pipe.out_host = [], [], []
            pipe.land_free, pipe.done = [], []
            for _ in range(2):
                land = _Captured()
                for name in self._names():
                    setattr(land, name, torch.empty_like(getattr(ent.st, name)))
                land.weight_list = getattr(land, "weight_list", None)
                land.weight_rel_list = getattr(land, "weight_rel_list", None)
                pipe.land.append(land)
                od = dict(cand_idx=torch.empty_like(cand_idx), pred_dist=torch.empty_like(pred_dist),
                          cand_count=torch.empty_like(cand_count), pred=torch.empty_like(pred),
                          loss=torch.empty_like(loss), status=torch.empty_like(db.graph.status))
                pipe.out_dev.append(od)
                pipe.out_host.append({k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                                      for k, v in od.items()})
                pipe.land_free.append(None)
                pipe.done.append(None)
                # pinned host staging of the inputs: pageable loader output (numpy, int64 / float64) is cast and copied
                # here by the host (memcpy speed), the DMA to the landing set then runs asynchronously
                stage = _Captured()
                for name in self._names():
                    t = getattr(ent.st, name)
                    setattr(stage, name, torch.empty(t.shape, dtype=t.dtype, pin_memory=True))
                stage.weight_list = getattr(stage, "weight_list", None)
                stage.weight_rel_list = getattr(stage, "weight_rel_list", None)
                pipe.stage = getattr(pipe, "stage", [])
                pipe.stage.append(stage)
                pipe.h2d_done = getattr(pipe, "h2d_done", [])
                pipe.h2d_done.append(None)
            ent.pipe = pipe
        return ent.pipe

    def submit(self, batch):
        """Enqueue one step (H2D of ``batch`` on the copy stream, graph replay, async D2H of the results) and
        return a :class:`Ticket`.  At most two tickets may be outstanding; ``collect`` them in order."""
        ent = self._entry(batch)
        pipe = self._pipe(ent)
        slot, self._slot = self._slot, self._slot ^ 1
        cs, cur = self._copy_stream, torch.cuda.current_stream()
        if pipe.done[slot] is not None:
            pipe.done[slot].synchronize()            # host buffers of this slot have been read out
        if pipe.land_free[slot] is not None:
            cs.wait_event(pipe.land_free[slot])
        land = pipe.land[slot]
        src = batch
        if self._needs_staging(batch):
            if pipe.h2d_done[slot] is not None:
                pipe.h2d_done[slot].synchronize()    # the previous DMA out of this staging set has finished
            src = self._stage_host(pipe.stage[slot], batch)
        with torch.cuda.stream(cs):
            self._fill(land, src)
            h2d_done = torch.cuda.Event()
            h2d_done.record(cs)
        pipe.h2d_done[slot] = h2d_done
        cur.wait_event(h2d_done)
        for name in self._names():                   # landing set -> the graph's static inputs (D2D, ~10 us)
            getattr(ent.st, name).copy_(getattr(land, name), non_blocking=True)
        pipe.land_free[slot] = torch.cuda.Event()
        pipe.land_free[slot].record(cur)
        ent.g.replay()
        db, loss, pred, pred_dist, cand_idx, cand_count, _ = ent.outs
        od = pipe.out_dev[slot]
        od["cand_idx"].copy_(cand_idx, non_blocking=True)
        od["pred_dist"].copy_(pred_dist, non_blocking=True)
        od["cand_count"].copy_(cand_count, non_blocking=True)
        od["pred"].copy_(pred, non_blocking=True)
        od["loss"].copy_(loss, non_blocking=True)
        od["status"].copy_(db.graph.status, non_blocking=True)
        out_ready = torch.cuda.Event()
        out_ready.record(cur)
        ds = self._d2h_stream
        ds.wait_event(out_ready)
        with torch.cuda.stream(ds):
            for k, v in od.items():
                pipe.out_host[slot][k].copy_(v, non_blocking=True)
            pipe.done[slot] = torch.cuda.Event()
            pipe.done[slot].record(ds)
        db.h2d_bytes = self._h2d_bytes
        self.model.last_batch = db
        t = Ticket()
        t.slot, t.done, t.ent = slot, pipe.done[slot], ent
        le = batch[0]
        t.local_entity_host = le.cpu().numpy() if isinstance(le, torch.Tensor) else np.asarray(le)
        t.B, t.N = db.B, db.N
        return t

    def collect(self, ticket):
        """Wait for a submitted step and return (retrieved, d2h_bytes, loss, pred): the ordered candidate lists
        of every question (like :func:`evaluate.retrieve`), the bytes read back, the loss and the argmax.
        Raises if the CSR build flagged node / relation ids outside the batch."""
        from .evaluate import Retrieved
        ticket.done.synchronize()
        h = ticket.ent.pipe.out_host[ticket.slot]
        if int(h["status"][0]) != 0:
            raise RuntimeError("fact list contains node/relation ids outside the batch (clamped)")
        idx_h, dist_h = h["cand_idx"].numpy(), h["pred_dist"].numpy()
        counts = h["cand_count"].numpy()
        le = ticket.local_entity_host
        res = []
        for b, c in enumerate(counts.tolist()):
            ix = idx_h[b, :c].astype(np.int64)
            res.append(Retrieved(ix, le[b, ix].astype(np.int64), dist_h[b, ix]))
        nbytes = sum(v.numel() * v.element_size() for v in h.values())
        return res, nbytes, float(h["loss"]), h["pred"].numpy().copy()

    def retrieve(self, out):
        """Ordered candidate lists of a :class:`StepOutput` (one D2H), like evaluate.retrieve."""
        from .evaluate import Retrieved
        counts_h = out.cand_count.cpu().numpy()
        self._check(out.db)
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
