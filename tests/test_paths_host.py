"""CPU: the host half of gnn_rag_b200.paths (entity / label maps, on-path pruning, the networkx-ordered enumeration of
all shortest paths) against networkx itself -- the library llm/src/utils/graph_utils.py:49-75 calls.  The hop distances
the product takes from the device BFS (tests/test_paths_gpu.py) are supplied by a plain host BFS here."""
import collections

import numpy as np
import pytest

nx = pytest.importorskip("networkx")

from gnn_rag_b200 import paths


class HostPathGraph(paths.PathGraph):
    """PathGraph whose distances come from a host BFS (no CUDA): same maps, same adjacency order."""

    def __init__(self, triples):
        ent, label = {}, {}
        heads = np.empty(len(triples), dtype=np.int64)
        tails = np.empty(len(triples), dtype=np.int64)
        for k, (h, r, t) in enumerate(triples):
            hi = ent.setdefault(h, len(ent))
            ti = ent.setdefault(t, len(ent))
            heads[k], tails[k] = hi, ti
            label[(hi, ti) if hi <= ti else (ti, hi)] = r.strip()
        self.ent2id, self.id2ent = ent, list(ent)
        self.heads, self.tails, self.label = heads, tails, label
        self.N = max(len(ent), 1)
        self.csr = object() if len(triples) else None
        self._nb = collections.defaultdict(list)
        for u, v in zip(heads.tolist(), tails.tolist()):
            self._nb[u].append(v)
            self._nb[v].append(u)

    def distances(self, nodes):
        out = np.full((len(nodes), self.N), -1, dtype=np.int32)
        for i, s in enumerate(nodes):
            out[i, s] = 0
            dq = collections.deque([s])
            while dq:
                u = dq.popleft()
                for v in self._nb[u]:
                    if out[i, v] < 0:
                        out[i, v] = out[i, u] + 1
                        dq.append(v)
        return out


def ref_paths(q_entity, a_entity, triples):                    # graph_utils.py:10-21,49-75
    G = nx.Graph()
    for h, r, t in triples:
        G.add_edge(h, t, relation=r.strip())
    out = []
    for h in q_entity:
        if h not in G:
            continue
        for t in a_entity:
            if t not in G:
                continue
            try:
                for p in nx.all_shortest_paths(G, h, t):
                    out.append(p)
            except Exception:                                   # noqa: BLE001 -- NetworkXNoPath is swallowed at :64-65
                pass
    return [[(p[i], G[p[i]][p[i + 1]]["relation"], p[i + 1]) for i in range(len(p) - 1)] for p in out]


def random_triples(seed, n_ent, n_tri, n_rel, components=1):
    rs = np.random.RandomState(seed)
    per = n_ent // components
    tri = []
    for _ in range(n_tri):
        c = rs.randint(components)
        a, b = rs.randint(per, size=2) + c * per
        tri.append(("m.%03d" % a, " rel.%d " % rs.randint(n_rel), "m.%03d" % b))
    return tri


@pytest.mark.parametrize("seed,n_ent,n_tri,components", [(1, 30, 60, 1), (2, 200, 500, 1), (3, 120, 150, 3),
                                                         (4, 12, 80, 1)])
def test_host_walk_equals_networkx_in_content_and_order(seed, n_ent, n_tri, components):
    tri = random_triples(seed, n_ent, n_tri, 7, components)
    g = HostPathGraph(tri)
    rs = np.random.RandomState(seed + 100)
    names = list(g.ent2id)
    q = [names[i] for i in rs.randint(len(names), size=3)] + ["m.absent"]
    a = [names[i] for i in rs.randint(len(names), size=6)] + [q[0]]          # incl. a zero-length path (source == target)
    got = paths.get_truth_paths(q, a, g)
    want = ref_paths(q, a, tri)
    assert got == want
    assert any(len(p) > 1 for p in want) or n_tri < 100


def test_last_triple_wins_the_relation_label_and_self_loops_are_kept():
    tri = [("a", "r1", "b"), ("b", "r2", "a"), ("b", "r3", "c"), ("c", "self", "c")]
    g = HostPathGraph(tri)
    assert paths.get_truth_paths(["a"], ["c"], g) == ref_paths(["a"], ["c"], tri) == [[("a", "r2", "b"), ("b", "r3", "c")]]
    assert paths.get_truth_paths([], ["c"], g) == [] and paths.get_truth_paths(["zz"], ["c"], g) == []
