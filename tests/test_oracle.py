"""CPU: pin oracle/kgqa_oracle.py against the golden vectors the unmodified reference produced."""
import numpy as np
import pytest
import torch

from golden_io import Golden, names
from oracle import kgqa_oracle as O


@pytest.mark.parametrize("name", names())
def test_forward_matches_reference(name):
    g = Golden(name)
    loss, pred, dist, trace = O.forward(g.sd, g.args, g.num_entity, g.num_word, g.batch,
                                        return_trace=True)
    ref = torch.from_numpy(g.out["pred_dist"])
    # same torch ops in the same order: expect (near) bit-equality; tolerance 1e-6 relative
    assert torch.allclose(dist, ref, rtol=1e-6, atol=1e-12), (dist - ref).abs().max()
    assert abs(float(loss) - float(g.out["loss"])) <= 1e-6 * max(1.0, abs(float(g.out["loss"])))
    assert pred.tolist() == g.out["pred"].tolist()
    assert np.allclose(trace["h_final"].numpy(), g.out["h_final"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", names())
def test_ranking_matches_reference(name):
    g = Golden(name)
    got = O.rank_candidates(g.batch[0], g.batch[1], g.out["pred_dist"], g.num_entity, g.args["eps"])
    ids, probs = g.cand_lists()
    assert [[c for _, c, _ in r] for r in got] == ids
    assert [[p for _, _, p in r] for r in got] == probs      # fp32 -> float64, exact


@pytest.mark.parametrize("name", names("rearev"))
def test_isolated_reason_layer(name):
    g = Golden(name)
    L = g.layer
    B, N = g.batch[0].shape
    mats = O.FactMats(g.batch[2], B, N, g.args["normalized_gnn"])
    pe = g.sd.get("reasoning.pos_emb0.weight") if g.args.get("pos_emb") else None
    pei = g.sd.get("reasoning.pos_emb_inv0.weight") if g.args.get("pos_emb") else None
    W, b = g.sd["reasoning.rel_linear0.weight"], g.sd["reasoning.rel_linear0.bias"]
    nb = O.reason_layer(mats, torch.from_numpy(L["dist"]), torch.from_numpy(L["ins"]),
                        torch.from_numpy(L["rel_features"]), W, b, False, pe)
    nbi = O.reason_layer(mats, torch.from_numpy(L["dist"]), torch.from_numpy(L["ins"]),
                         torch.from_numpy(L["rel_features_inv"]), W, b, True, pei)
    D = L["ins"].shape[1]
    assert np.allclose(nb.view(B, N, D).numpy(), L["neighbor_rep"], rtol=1e-6, atol=1e-9)
    assert np.allclose(nbi.view(B, N, D).numpy(), L["neighbor_rep_inv"], rtol=1e-6, atol=1e-9)


def test_twin_nodes_tie_exactly():
    """Structural twins (local 4 and 5) must come out with identical floats in the reference."""
    g = Golden("rearev_sharp_ties")
    pd = g.out["pred_dist"]
    assert (pd[:, 4] == pd[:, 5]).all()


def test_shortest_path_nodes_vs_networkx():
    nx = pytest.importorskip("networkx")
    rs = np.random.RandomState(3)
    n = 60
    heads = rs.randint(0, n, size=140)
    tails = rs.randint(0, n, size=140)
    G = nx.Graph()
    for h, t in zip(heads.tolist(), tails.tolist()):
        G.add_edge(h, t)
    sources, targets = [0, 1], [7, 13, 22]
    want = set()
    for s in sources:                     # llm/src/utils/graph_utils.py:49-64
        if s not in G:
            continue
        for t in targets:
            if t not in G:
                continue
            try:
                for p in nx.all_shortest_paths(G, s, t):
                    want.update(p)
            except nx.NetworkXNoPath:
                pass
    got, _ = O.shortest_path_nodes(heads.tolist(), tails.tolist(), n, sources, targets)
    assert set(got) == want
