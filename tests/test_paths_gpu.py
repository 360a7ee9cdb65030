"""Relation-labelled shortest paths (gnn_rag_b200.paths) against the reference's own procedure -- networkx, the
library ``llm/src/utils/graph_utils.py`` calls -- restated here line by line: same paths, same labels, same ORDER."""
import networkx as nx
import numpy as np
import pytest

from gnn_rag_b200 import paths

pytestmark = pytest.mark.gpu


def ref_build_graph(graph):                                   # graph_utils.py:10-21 (encrypt=False)
    G = nx.Graph()
    for h, r, t in graph:
        G.add_edge(h, t, relation=r.strip())
    return G


def ref_get_truth_paths(q_entity, a_entity, graph):            # graph_utils.py:49-75
    out = []
    for h in q_entity:
        if h not in graph:
            continue
        for t in a_entity:
            if t not in graph:
                continue
            try:
                for p in nx.all_shortest_paths(graph, h, t):
                    out.append(p)
            except Exception:                                   # noqa: BLE001 -- the reference swallows NetworkXNoPath
                pass
    return [[(p[i], graph[p[i]][p[i + 1]]["relation"], p[i + 1]) for i in range(len(p) - 1)] for p in out]


def random_triples(seed, n_ent, n_tri, n_rel, components=1):
    rs = np.random.RandomState(seed)
    tri = []
    per = n_ent // components
    for _ in range(n_tri):
        c = rs.randint(components)
        a, b = rs.randint(per, size=2) + c * per
        tri.append(("m.%03d" % a, " rel.%d " % rs.randint(n_rel), "m.%03d" % b))    # labels are strip()ped
    return tri


@pytest.mark.parametrize("seed,n_ent,n_tri,components", [(1, 30, 60, 1), (2, 200, 500, 1), (3, 120, 150, 3),
                                                         (4, 12, 80, 1), (5, 2000, 6000, 1)])
def test_truth_paths_equal_networkx_in_content_and_order(seed, n_ent, n_tri, components):
    tri = random_triples(seed, n_ent, n_tri, 7, components)
    tri += [tri[0], (tri[1][2], " other ", tri[1][0]), (tri[2][0], "loop", tri[2][0])]   # duplicate, reversed relabel, self loop
    rs = np.random.RandomState(100 + seed)
    ents = sorted({h for h, _, _ in tri} | {t for _, _, t in tri})
    q = list(rs.choice(ents, size=2, replace=False)) + ["m.not_in_graph"]
    a = list(rs.choice(ents, size=min(8, len(ents)), replace=False)) + [q[0], "m.absent"]
    want = ref_get_truth_paths(q, a, ref_build_graph(tri))
    got = paths.get_truth_paths(q, a, paths.build_graph(tri))
    assert got == want
    assert any(len(p) == 0 for p in got)                        # h == t gives the empty path, as in the reference


def test_truth_paths_empty_inputs():
    g = paths.build_graph([])
    assert paths.get_truth_paths(["a"], ["b"], g) == []
    g = paths.build_graph([("a", "r", "b")])
    assert paths.get_truth_paths([], ["b"], g) == [] and paths.get_truth_paths(["a"], [], g) == []
    assert paths.get_truth_paths(["a"], ["b"], g) == [[("a", "r", "b")]]
