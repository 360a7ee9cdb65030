"""CPU: host-side logic -- checkpoint (state_dict) compatibility with the reference, synthetic batch
invariants, question sharding + gloo all-gather (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gnn_rag_b200 as G
from gnn_rag_b200 import parallel, synthetic as S
from golden_io import Golden, names


@pytest.mark.parametrize("name", names(include_lm=True))
def test_reference_state_dict_loads_strict(name):
    g = Golden(name)
    cls = G.NSM if g.args["model_name"] == "NSM" else G.ReaRev
    m = cls(dict(g.args), g.num_entity, g.num_relation, g.num_word)
    ours = m.state_dict()
    assert set(ours) == set(g.sd), set(ours) ^ set(g.sd)
    for k, v in g.sd.items():
        assert tuple(ours[k].shape) == tuple(v.shape), k
    m.load_state_dict(g.sd, strict=True)
    # and the other way round: a checkpoint written by us has the reference layout
    ck = {"model_state_dict": m.state_dict()}
    assert list(ck["model_state_dict"].keys())


def test_synthetic_batch_layout():
    b = S.make_batch(3, B=4, N=30, E=80, num_entity=500, num_relation=12, num_word=50, n_real="ragged",
                     multi_seed=True, test=True)
    le, qe, kb, qi, sd, tb, ad, al = b
    heads, rels, tails, bids, fids, wl, wrl = kb
    assert le.dtype == np.int64 and qe.dtype == np.float64 and tb is None
    assert np.allclose(sd.sum(1), 1.0)
    assert (heads // 30 == bids).all() and (tails // 30 == bids).all()       # block diagonal
    assert rels.max() == 11 and (rels[heads == tails].max() == 11)             # self loops use R-1
    assert len(wl) == len(heads) == len(wrl)
    real = le != 500
    for q in range(4):                                                          # one self loop per real node
        n_self = ((rels == 11) & (bids == q)).sum()
        assert n_self == real[q].sum()
    assert (fids == np.arange(len(heads))).all()


def test_question_range_is_a_partition():
    for B in (1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            spans = [parallel.question_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_reassembles():
    b = S.make_batch(5, B=5, N=20, E=50, num_entity=300, num_relation=9, num_word=40)
    N = 20
    parts = [parallel.shard_batch(b, r, 2) for r in range(2)]
    assert sum(p[0].shape[0] for p in parts) == 5
    off = 0
    got = []
    for p in parts:
        h, r, t = p[2][0], p[2][1], p[2][2]
        assert h.min() >= 0 and h.max() < p[0].shape[0] * N
        got.append(np.stack([h + off * N, r, t + off * N], 1))
        off += p[0].shape[0]
    got = np.concatenate(got)
    want = np.stack([b[2][0], b[2][1], b[2][2]], 1)
    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist()))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, B, N, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(B * N, dtype=torch.float32).view(B, N)
    lo, hi = parallel.question_range(B, rank, world)
    out = parallel.all_gather_scores(full[lo:hi].clone(), B)
    q.put((rank, bool(torch.equal(out, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 7])
def test_all_gather_scores_gloo_world2(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, B, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's reference arm) runs without a GPU and prints ONE JSON line with the
    contract keys; it times the oracle port of the reference op sequence on the host cores."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] and line["unit"] == "questions/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"]


def _reference_merge(cand1, cand2):
    """The candidate-union loop of load_gnn_rag, llm/src/qa_prediction/predict_answer.py:61-75, restated."""
    cand1 = [list(c) for c in cand1]
    for c2 in cand2:
        found = False
        for c1 in cand1:
            if c2[0] == c1[0]:
                if c2[1] > c1[1]:
                    c1[1] = c2[1]
                found = True
                break
        if not found:
            cand1.append(list(c2))
    return sorted(cand1, key=lambda x: x[1], reverse=True)


def test_merge_candidates_matches_the_llm_stage_union():
    from gnn_rag_b200 import evaluate
    rs = np.random.RandomState(0)
    for _ in range(50):
        ents = ["m.%d" % e for e in rs.randint(0, 12, 20)]
        n1, n2 = rs.randint(0, 9), rs.randint(0, 9)
        # scores on a coarse grid -> exact ties, duplicates inside a list -> first-match semantics
        c1 = [[ents[i], float(rs.randint(0, 5)) / 4] for i in range(n1)]
        c2 = [[ents[10 + i], float(rs.randint(0, 5)) / 4] for i in range(n2)]
        keep1, keep2 = [list(c) for c in c1], [list(c) for c in c2]
        assert evaluate.merge_candidates(c1, c2) == _reference_merge(c1, c2)
        assert c1 == keep1 and c2 == keep2                        # inputs untouched
    rows1 = [{"question": "q", "cand": [["a", 0.6], ["b", 0.3]]}]
    rows2 = [{"question": "q", "cand": [["b", 0.5], ["c", 0.4]]}]
    assert evaluate.merge_info_rows(rows1, rows2)[0]["cand"] == [["a", 0.6], ["b", 0.5], ["c", 0.4]]
