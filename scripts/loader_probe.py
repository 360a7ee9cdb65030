"""Host-side timing of the batch assembly at BASELINE cfg2 scale (64 questions x 6000 KG facts, 2000 local entities,
6107 relation rows): the unmodified reference `BasicDataLoader._build_fact_mat` (gnn/dataset_load.py:473-527) vs
gnn_rag_b200.loader.build_fact_mat, same RNG state, outputs compared bit for bit.  CPU only; needs /root/reference."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gnn_rag_b200 import loader  # noqa: E402
from oracle import ref_harness  # noqa: E402


class Loader:
    data_eff = False
    use_self_loop = True

    def __init__(self, B, N, E, R1, seed=0):
        rs = np.random.RandomState(seed)
        self.max_local_entity, self.num_kb_relation = N, R1
        self.kb_adj_mats = [(rs.randint(0, N, E).astype(int), rs.randint(0, R1 - 1, E).astype(int),
                             rs.randint(0, N, E).astype(int)) for _ in range(B)]
        self.global2local_entity_maps = [dict.fromkeys(range(N)) for _ in range(B)]


def timeit(fn, n):
    best = 1e9
    for i in range(n):
        np.random.seed(i)
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


if __name__ == "__main__":
    B, N, E, R1 = 64, 2000, 6000, 6107
    ld = Loader(B, N, E, R1)
    ids = list(range(B))
    res = {"config": {"B": B, "N": N, "E": E, "R1": R1, "facts_incl_self_loops": B * (E + N)}}
    t_new, got = timeit(lambda: loader.build_fact_mat(ld, ids, 0.0), 5)
    res["ours_lists_ms"] = t_new * 1e3
    res["ours_arrays_ms"] = timeit(lambda: loader.build_fact_mat(ld, ids, 0.0, weights="arrays"), 5)[0] * 1e3
    res["ours_no_weights_int32_ms"] = timeit(
        lambda: loader.build_fact_mat(ld, ids, 0.0, weights="none", index_dtype=np.int32), 5)[0] * 1e3
    loader.preconvert(ld)
    res["ours_unshuffled_no_weights_int32_ms"] = timeit(
        lambda: loader.build_fact_mat(ld, ids, 0.0, weights="none", index_dtype=np.int32, shuffle=False), 5)[0] * 1e3
    if ref_harness.available():
        ref_harness._import_reference()
        import dataset_load
        t_ref, want = timeit(lambda: dataset_load.BasicDataLoader._build_fact_mat(ld, ids, 0.0), 2)
        res["reference_ms"] = t_ref * 1e3
        np.random.seed(1); a = loader.build_fact_mat(ld, ids, 0.0)
        np.random.seed(1); b = dataset_load.BasicDataLoader._build_fact_mat(ld, ids, 0.0)
        res["bit_identical"] = all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(a, b))
        res["speedup_lists"] = res["reference_ms"] / res["ours_lists_ms"]
    res["host"] = {"cpus": os.cpu_count(), "threads_used": 1}
    print(json.dumps(res, indent=1))
