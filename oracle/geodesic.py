"""CPU restatement of the geodesic GraphPOPE path (reference utils.py:64-147).

TEST INFRASTRUCTURE — see oracle/__init__.py.  Three tiers:

T0  ``t0_pairwise``      utils.py:64-81 restated: N*K ``nx.shortest_path`` calls
                         (networkx bidirectional BFS), ``1/len(path)``, ``0`` when
                         there is no path.  The reference algorithm itself; slow.
T1  ``t1_sssp_reverse``  K x ``nx.single_source_shortest_path_length`` on the
                         reversed DiGraph; bit-equal to T0 (tests pin this).
T2  ``t2_bfs_csr``       numpy frontier BFS from each anchor over the in-edge
                         lists (i.e. over reversed edges); equal to T1.
The plain-C tier lives in bfs_oracle.c / cbfs.py.

Conventions reproduced (all from utils.py):
  * graph = ``to_networkx(data)`` defaults -> DiGraph, parallel edges collapse,
    self-loops kept, no symmetrisation (utils.py:121);
  * row = source node, column = target anchor, value = 1/len(path) = 1/(d+1)
    with d = hops(node -> anchor); node == anchor -> 1.0; unreachable -> 0
    (utils.py:72-76);
  * float64 division rounded to float32 by ``torch.as_tensor`` (utils.py:125) —
    identical to IEEE fp32 ``1.0f/(float)(d+1)`` for every d+1 in [1, 65536]
    (checked in tests/test_oracle_geodesic.py).
"""
from __future__ import annotations

import numpy as np

UNREACHABLE_U16 = 0xFFFF


# --------------------------------------------------------------------------- graph
def dedup_edges(edge_index: np.ndarray, num_nodes: int, symmetrize: bool = False):
    """Unique directed edges sorted by (src, dst) — what ``to_networkx`` keeps.

    Restates PyG ``to_networkx`` as used at utils.py:121: ``add_edge(u, v)`` per
    column collapses duplicates; self-loops stay.
    """
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    src, dst = ei[0], ei[1]
    if src.size and (src.min() < 0 or dst.min() < 0 or src.max() >= num_nodes
                     or dst.max() >= num_nodes):
        raise ValueError("edge_index entry outside [0, num_nodes)")
    if symmetrize:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    code = np.unique(src * np.int64(max(num_nodes, 1)) + dst)
    n = np.int64(max(num_nodes, 1))
    return (code // n).astype(np.int64), (code % n).astype(np.int64)


def csr_from_sorted(rows: np.ndarray, cols: np.ndarray, num_nodes: int):
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    np.cumsum(rowptr, out=rowptr)
    return rowptr, cols.astype(np.int64)


def out_csr(edge_index, num_nodes, symmetrize=False):
    """CSR by source: row u lists the out-neighbours of u (sorted, unique)."""
    s, d = dedup_edges(edge_index, num_nodes, symmetrize)
    return csr_from_sorted(s, d, num_nodes)


def in_csr(edge_index, num_nodes, symmetrize=False):
    """CSR by destination: row v lists the in-neighbours of v (sorted, unique)."""
    s, d = dedup_edges(edge_index, num_nodes, symmetrize)
    order = np.lexsort((s, d))
    return csr_from_sorted(d[order], s[order], num_nodes)


def to_digraph(edge_index, num_nodes):
    """networkx DiGraph exactly as the ``to_networkx`` default builds it."""
    import networkx as nx

    G = nx.DiGraph()
    G.add_nodes_from(range(int(num_nodes)))
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    G.add_edges_from(zip(ei[0].tolist(), ei[1].tolist()))
    return G


# --------------------------------------------------------------------------- values
def normalise_hops(dist_u16: np.ndarray) -> np.ndarray:
    """uint16 hop matrix -> float32 features, utils.py:73,76 convention."""
    d = np.asarray(dist_u16)
    out = np.zeros(d.shape, dtype=np.float32)
    reach = d != UNREACHABLE_U16
    # float64 1/(d+1) rounded to float32, as torch.as_tensor does (utils.py:125)
    out[reach] = (1.0 / (d[reach].astype(np.float64) + 1.0)).astype(np.float32)
    return out


# --------------------------------------------------------------------------- T0
def t0_rows(G, anchors, nodes):
    """utils.py:64-81 (``shortest_path_length``) restated; returns {node: [K]}."""
    import networkx as nx

    out = {}
    for node in nodes:
        vals = []
        for a in anchors:
            try:
                vals.append(1 / len(nx.shortest_path(G, source=node, target=a)))
            except nx.NetworkXNoPath:
                vals.append(0)
        out[node] = vals
    return out


def t0_pairwise(edge_index, num_nodes, anchors) -> np.ndarray:
    G = to_digraph(edge_index, num_nodes)
    rows = t0_rows(G, [int(a) for a in anchors], range(num_nodes))
    k = len(anchors)
    if num_nodes == 0 or k == 0:
        return np.zeros((num_nodes, k), dtype=np.float32)
    return np.asarray([rows[i] for i in range(num_nodes)], dtype=np.float64).astype(np.float32)


# --------------------------------------------------------------------------- T1
def t1_sssp_reverse_hops(edge_index, num_nodes, anchors) -> np.ndarray:
    import networkx as nx

    G = to_digraph(edge_index, num_nodes).reverse(copy=False)
    k = len(anchors)
    dist = np.full((num_nodes, k), UNREACHABLE_U16, dtype=np.uint16)
    for j, a in enumerate(anchors):
        for node, d in nx.single_source_shortest_path_length(G, int(a)).items():
            dist[node, j] = d
    return dist


# --------------------------------------------------------------------------- T2
def t2_bfs_csr_hops(edge_index, num_nodes, anchors, symmetrize=False) -> np.ndarray:
    """Frontier BFS per anchor over in-edge lists; returns uint16 ``[N, K]``."""
    rowptr, col = in_csr(edge_index, num_nodes, symmetrize)
    k = len(anchors)
    dist = np.full((num_nodes, k), UNREACHABLE_U16, dtype=np.uint16)
    for j, a in enumerate(anchors):
        a = int(a)
        seen = np.zeros(num_nodes, dtype=bool)
        seen[a] = True
        dist[a, j] = 0
        frontier = np.array([a], dtype=np.int64)
        level = 0
        while frontier.size:
            level += 1
            if level >= UNREACHABLE_U16:
                raise OverflowError("hop distance does not fit uint16")
            starts = rowptr[frontier]
            counts = rowptr[frontier + 1] - starts
            if counts.sum() == 0:
                break
            offs = np.repeat(starts - np.concatenate([[0], np.cumsum(counts)[:-1]]), counts)
            nbrs = col[offs + np.arange(counts.sum())]
            nbrs = np.unique(nbrs[~seen[nbrs]])
            seen[nbrs] = True
            dist[nbrs, j] = level
            frontier = nbrs
    return dist


def geodesic_features(edge_index, num_nodes, anchors, symmetrize=False) -> np.ndarray:
    """float32 ``[N, K]`` block of utils.py:116-126, via the T2 tier."""
    return normalise_hops(t2_bfs_csr_hops(edge_index, num_nodes, anchors, symmetrize))


def concat_features(x: np.ndarray, emb: np.ndarray) -> np.ndarray:
    """utils.py:129-135: ``torch.cat((data.x, embedding), 1)``."""
    return np.concatenate([np.asarray(x, dtype=np.float32),
                           np.asarray(emb, dtype=np.float32)], axis=1)


# --------------------------------------------------------------------------- pool split
def node_slices(num_nodes: int, num_workers: int):
    """utils.py:100: float-arithmetic slice bounds ``int(N/w*i)``."""
    return [(int(num_nodes / num_workers * i), int(num_nodes / num_workers * (i + 1)))
            for i in range(num_workers)]
