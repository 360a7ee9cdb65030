"""TEST INFRASTRUCTURE -- CPU restatement of the reference's batch assembly, ``BasicDataLoader._build_fact_mat``
(gnn/dataset_load.py:473-527), kept in the reference's own form (four ``np.append`` per question, two Python
``Counter`` passes over all facts) so that timing it stands for the reference's ``get_batch`` cost in bench.py's
``--impl reference`` / ``cpu_baseline`` legs.  Pinned against arrays recorded from the unmodified reference
(tests/golden/loader/*.npz, tests/test_loader.py).  Not part of the product (the product's drop-in is
gnn_rag_b200/loader.py)."""
from collections import Counter

import numpy as np


class LoaderState:
    """The attributes ``_build_fact_mat`` reads (dataset_load.py:96-118, 262-335)."""

    def __init__(self, kb_adj_mats, num_entities, max_local_entity, num_kb_relation, use_self_loop=True):
        self.kb_adj_mats = kb_adj_mats
        self.global2local_entity_maps = [range(n) for n in num_entities]     # only len() is used (:500)
        self.max_local_entity = max_local_entity
        self.num_kb_relation = num_kb_relation
        self.use_self_loop = use_self_loop
        self.data_eff = False


def build_fact_mat(self, sample_ids, fact_dropout):
    batch_heads = np.array([], dtype=int)                                    # :477-480
    batch_rels = np.array([], dtype=int)
    batch_tails = np.array([], dtype=int)
    batch_ids = np.array([], dtype=int)
    for i, sample_id in enumerate(sample_ids):                               # :482
        index_bias = i * self.max_local_entity
        head_list, rel_list, tail_list = self.kb_adj_mats[sample_id]        # :487
        num_fact = len(head_list)
        num_keep_fact = int(np.floor(num_fact * (1 - fact_dropout)))         # :489
        mask_index = np.random.permutation(num_fact)[: num_keep_fact]        # :490
        batch_heads = np.append(batch_heads, head_list[mask_index] + index_bias)   # :492-498
        batch_rels = np.append(batch_rels, rel_list[mask_index])
        batch_tails = np.append(batch_tails, tail_list[mask_index] + index_bias)
        batch_ids = np.append(batch_ids, np.full(len(mask_index), i, dtype=int))
        if self.use_self_loop:                                               # :499-506
            num_ent_now = len(self.global2local_entity_maps[sample_id])
            ent_array = np.array(range(num_ent_now), dtype=int) + index_bias
            rel_array = np.array([self.num_kb_relation - 1] * num_ent_now, dtype=int)
            batch_heads = np.append(batch_heads, ent_array)
            batch_tails = np.append(batch_tails, ent_array)
            batch_rels = np.append(batch_rels, rel_array)
            batch_ids = np.append(batch_ids, np.full(num_ent_now, i, dtype=int))
    fact_ids = np.array(range(len(batch_heads)), dtype=int)                  # :507
    head_count = Counter(batch_heads)                                        # :509
    weight_list = [1.0 / head_count[head] for head in batch_heads]           # :511
    head_rels_batch = list(zip(batch_heads, batch_rels))                     # :514
    head_rels_count = Counter(head_rels_batch)
    weight_rel_list = [1.0 / head_rels_count[(h, r)] for (h, r) in head_rels_batch]   # :517
    return batch_heads, batch_rels, batch_tails, batch_ids, fact_ids, weight_list, weight_rel_list


def state_from_batch(batch, num_kb_relation):
    """Per-question (head, rel, tail) arrays of a synthetic ``get_batch`` tuple (gnn_rag_b200.synthetic.make_batch), i.e.
    what the loader held before batching: local node ids, self loops removed."""
    le = batch[0]
    heads, rels, tails, bids = batch[2][:4]
    B, N = le.shape
    mats, nents = [], []
    order = np.argsort(bids, kind="stable")
    cuts = np.searchsorted(bids[order], np.arange(B + 1))
    for b in range(B):
        sel = order[cuts[b]:cuts[b + 1]]
        keep = sel[rels[sel] != num_kb_relation - 1]
        mats.append((heads[keep] - b * N, rels[keep].copy(), tails[keep] - b * N))
        loops = sel[rels[sel] == num_kb_relation - 1]
        nents.append(len(loops))
    return LoaderState(mats, nents, N, num_kb_relation)
