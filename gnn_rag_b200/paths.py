"""Relation-labelled shortest paths between question entities and retrieved candidates -- the hand-over of the GNN
stage to the LLM stage (``llm/src/utils/graph_utils.py:10-21,49-75`` consumed by
``llm/src/qa_prediction/build_qa_input.py:114-127``).

    from gnn_rag_b200.paths import build_graph, get_truth_paths          # same names / arguments as utils.*
    graph = build_graph(question_dict["graph"])
    reasoning_paths = get_truth_paths(question_dict["q_entity"], question_dict["cand"], graph)

Same result as the reference, INCLUDING the order of the paths (the prompt text lists them in that order): the
reference builds an undirected ``networkx.Graph`` (a later triple between the same two entities overwrites the
relation label, either direction) and concatenates ``nx.all_shortest_paths(graph, h, t)`` over all (h, t) pairs.
Here the breadth-first distances from every question entity and every candidate come from the device kernel
(csrc/paths.cu, level-synchronous BFS over both CSRs); a node lies on a shortest h-t path iff d(h,v) + d(v,t) = d(h,t).
Only that sub-graph (a handful of nodes) is then walked on the host, in networkx's order: its BFS predecessor lists
depend only on the insertion order of the on-path nodes' edges -- every predecessor of an on-path node is itself
on-path -- so pruning does not change the order in which paths come out.
"""
import numpy as np
import torch

from . import ops


class PathGraph:
    """Device-resident undirected adjacency of one question's triple list + the host-side label maps."""

    def __init__(self, triples, device=None):
        device = torch.device("cuda") if device is None else torch.device(device)
        ent = {}
        heads = np.empty(len(triples), dtype=np.int64)
        tails = np.empty(len(triples), dtype=np.int64)
        label = {}
        for k, (h, r, t) in enumerate(triples):                # node / edge insertion order of nx.Graph.add_edge
            hi = ent.setdefault(h, len(ent))
            ti = ent.setdefault(t, len(ent))
            heads[k], tails[k] = hi, ti
            label[(hi, ti) if hi <= ti else (ti, hi)] = r.strip()      # graph_utils.py:20 -- the last triple wins
        self.ent2id, self.id2ent = ent, list(ent)
        self.heads, self.tails, self.label = heads, tails, label
        self.N = max(len(ent), 1)
        self.device = device
        if len(triples):
            z = torch.zeros(len(triples), dtype=torch.int64, device=device)
            self.csr = ops.csr_build(torch.from_numpy(heads).to(device), z, torch.from_numpy(tails).to(device), 1, self.N, 1)
        else:
            self.csr = None

    def __contains__(self, entity):
        return entity in self.ent2id

    def distances(self, nodes):
        """BFS hop distances from each of ``nodes`` (ids) to every node: int32 [len(nodes), N], -1 = unreachable."""
        k = len(nodes)
        src = torch.tensor([nodes], dtype=torch.int32, device=self.device)
        cnt = torch.tensor([k], dtype=torch.int32, device=self.device)
        one = torch.zeros(1, 1, dtype=torch.int32, device=self.device)
        _on, _pd, dist = ops.shortest_path_nodes(self.csr, src, cnt, one, torch.zeros(1, dtype=torch.int32, device=self.device),
                                                 return_distances=True)
        return dist[0, :k].cpu().numpy()


def build_graph(graph, entities=None, encrypt=False, device=None):
    """Drop-in for ``utils.build_graph`` (graph_utils.py:10-21); ``encrypt`` re-labels entities through the reference's
    ``entities_names.json`` and is not supported here."""
    if encrypt:
        raise NotImplementedError("encrypt=True needs the reference's entities_names.json (graph_utils.py:6-8)")
    return PathGraph(graph, device)


def _ordered_adjacency(g, on):
    """Neighbour lists of the on-path nodes in networkx's insertion order, restricted to on-path nodes."""
    sel = np.nonzero(on[g.heads] & on[g.tails])[0]
    adj = {}
    for k in sel.tolist():
        u, v = int(g.heads[k]), int(g.tails[k])
        adj.setdefault(u, {}).setdefault(v, None)
        adj.setdefault(v, {}).setdefault(u, None)
    return adj


def _all_shortest_paths(adj, source, target):
    """networkx.all_shortest_paths on an unweighted graph: ``predecessor`` (unweighted.py) + the stack walk of
    ``_build_paths_from_predecessors`` (generic.py), order preserved."""
    seen, pred, nextlevel, level = {source: 0}, {source: []}, [source], 0
    while nextlevel:
        level += 1
        thislevel, nextlevel = nextlevel, []
        for v in thislevel:
            for w in adj.get(v, ()):
                if w not in seen:
                    pred[w] = [v]
                    seen[w] = level
                    nextlevel.append(w)
                elif seen[w] == level:
                    pred[w].append(v)
    if target not in pred:
        return
    on_stack = {target}
    stack, top = [[target, 0]], 0
    while top >= 0:
        node, i = stack[top]
        if node == source:
            yield [p for p, _ in reversed(stack[: top + 1])]
        if len(pred[node]) > i:
            stack[top][1] = i + 1
            nxt = pred[node][i]
            if nxt in on_stack:
                continue
            on_stack.add(nxt)
            top += 1
            if top == len(stack):
                stack.append([nxt, 0])
            else:
                stack[top][:] = [nxt, 0]
        else:
            on_stack.discard(node)
            top -= 1


def get_truth_paths(q_entity, a_entity, graph):
    """Drop-in for ``utils.get_truth_paths`` (graph_utils.py:49-75): every shortest path between every question entity
    and every candidate, as ``[(u, relation, v), ...]`` triples, in the reference's order."""
    g = graph
    hs = [h for h in q_entity if h in g]
    ts = [t for t in a_entity if t in g]
    if not hs or not ts or g.csr is None:
        return []
    uniq = list(dict.fromkeys(hs + ts))
    dist = g.distances([g.ent2id[e] for e in uniq])
    row = {e: i for i, e in enumerate(uniq)}
    out = []
    for h in hs:
        dh = dist[row[h]]
        hi = g.ent2id[h]
        for t in ts:
            ti = g.ent2id[t]
            d = int(dh[ti])
            if d < 0:
                continue                                       # nx.NetworkXNoPath, swallowed at :64-65
            dt = dist[row[t]]
            on = (dh >= 0) & (dt >= 0) & (dh + dt == d)
            adj = _ordered_adjacency(g, on)
            for p in _all_shortest_paths(adj, hi, ti):
                out.append([(g.id2ent[p[i]],
                             g.label[(p[i], p[i + 1]) if p[i] <= p[i + 1] else (p[i + 1], p[i])],
                             g.id2ent[p[i + 1]]) for i in range(len(p) - 1)])
    return out
