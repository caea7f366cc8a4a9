"""CPU restatement of the anchor samplers (reference utils.py:18-62).

TEST INFRASTRUCTURE — see oracle/__init__.py.

Only the samplers the device path re-implements are restated here
(stochastic, degree_centrality, pagerank, closeness_centrality, clustering_coefficient).
The other two centralities stay on the reference's own networkx calls in the product
(north_star).

Third-party arithmetic restated: networkx (unpinned by the reference's
requirements.txt; 3.6.1 installed) ``degree_centrality`` and
``_pagerank_scipy`` — utils.py:28,40 call sites.
"""
from __future__ import annotations

import numpy as np

from .geodesic import dedup_edges


def stochastic(num_nodes: int, k: int) -> np.ndarray:
    """utils.py:22-24: draws WITH replacement from the global legacy numpy RNG."""
    return np.random.choice(np.arange(num_nodes), k)


def stable_top_k(score: np.ndarray, k: int) -> list:
    """utils.py:29-30 pattern: stable ascending sort by score, keep the last k.

    Ties keep ascending node id, so the cut favours larger ids; the returned
    list is in ascending-score order.  ``k == 0`` returns ALL nodes because
    ``list[-0:]`` is the whole list (reference quirk).
    """
    order = np.argsort(score, kind="stable")
    return [int(i) for i in order.tolist()[-k:]]


def degree_scores(edge_index, num_nodes: int) -> np.ndarray:
    """In+out degree over de-duplicated directed edges (self-loop counts 2).

    ``nx.degree_centrality`` multiplies by 1/(N-1), a monotone map, so the
    integer degree orders identically (utils.py:40-42).
    """
    s, d = dedup_edges(edge_index, num_nodes)
    deg = np.zeros(num_nodes, dtype=np.int64)
    np.add.at(deg, s, 1)
    np.add.at(deg, d, 1)
    return deg


def degree_centrality_anchors(edge_index, num_nodes: int, k: int) -> list:
    return stable_top_k(degree_scores(edge_index, num_nodes), k)


def pagerank_scores(edge_index, num_nodes: int, alpha: float = 0.85,
                    tol: float = 1.0e-6, max_iter: int = 100):
    """networkx ``_pagerank_scipy`` defaults restated operation by operation.

    Returns ``(x, iterations)``.  Summation order follows scipy's
    ``csc_matvec`` used for ``x @ A``: y[v] accumulates a_uv * x[u] over the
    in-neighbours u of v in ascending u, multiply then add (no FMA); the
    dangling mass is Python's left-to-right ``sum``.
    """
    n = int(num_nodes)
    if n == 0:
        return np.zeros(0), 0
    s, d = dedup_edges(edge_index, n)
    outdeg = np.zeros(n, dtype=np.float64)
    np.add.at(outdeg, s, 1.0)
    inv = np.zeros(n, dtype=np.float64)
    nz = outdeg != 0
    inv[nz] = 1.0 / outdeg[nz]
    # in-edge lists, ascending source within each destination
    order = np.lexsort((s, d))
    src_by_dst = s[order]
    dst_sorted = d[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, dst_sorted + 1, 1)
    np.cumsum(rowptr, out=rowptr)
    dangling = np.where(~nz)[0]
    x = np.repeat(1.0 / n, n)
    p = np.repeat(1.0 / n, n)
    maxdeg = int((rowptr[1:] - rowptr[:-1]).max()) if n else 0
    for it in range(1, max_iter + 1):
        xlast = x
        contrib = inv[src_by_dst] * x[src_by_dst]
        # sequential (left-to-right) accumulation per destination, vectorised
        # across destinations: add the t-th in-neighbour of every row at step t
        y = np.zeros(n, dtype=np.float64)
        deg = rowptr[1:] - rowptr[:-1]
        for t in range(maxdeg):
            rows = np.where(deg > t)[0]
            y[rows] = y[rows] + contrib[rowptr[rows] + t]
        dsum = 0
        for v in x[dangling]:
            dsum = dsum + v
        x = alpha * (y + dsum * p) + (1 - alpha) * p
        err = np.absolute(x - xlast).sum()
        if err < n * tol:
            return x, it
    raise RuntimeError("pagerank: power iteration failed to converge in %d iterations"
                       % max_iter)


def pagerank_anchors(edge_index, num_nodes: int, k: int) -> list:
    x, _ = pagerank_scores(edge_index, num_nodes)
    return stable_top_k(x, k)


def closeness_scores(edge_index, num_nodes: int) -> np.ndarray:
    """networkx ``closeness_centrality(G)`` defaults restated (utils.py:50-54 call site).

    Incoming distance on the DiGraph, Wasserman-Faust scaling, float64 in networkx's
    operation order: ``cc = (r - 1) / totsp; cc *= (r - 1) / (N - 1)`` with ``r`` the
    number of nodes that can reach the node (itself included) and ``totsp`` the sum of
    their hop counts; 0 when ``totsp == 0`` or ``N == 1``.
    """
    from . import cbfs

    n = int(num_nodes)
    out = np.zeros(n, dtype=np.float64)
    if n == 0:
        return out
    csr = cbfs.InCsr(edge_index, n)
    step = 512
    for c0 in range(0, n, step):
        anchors = np.arange(c0, min(n, c0 + step), dtype=np.int64)
        hops = cbfs.bfs_hops(csr, anchors).astype(np.int64)  # [N, k], 0xFFFF = cannot reach
        reach = hops != 0xFFFF
        r = reach.sum(axis=0)
        totsp = np.where(reach, hops, 0).sum(axis=0)
        for j in range(anchors.size):
            if totsp[j] > 0 and n > 1:
                cc = (float(r[j]) - 1.0) / float(totsp[j])
                cc *= (float(r[j]) - 1.0) / (n - 1)
                out[c0 + j] = cc
    return out


def closeness_centrality_anchors(edge_index, num_nodes: int, k: int) -> list:
    return stable_top_k(closeness_scores(edge_index, num_nodes), k)


def clustering_scores(edge_index, num_nodes: int) -> np.ndarray:
    """networkx ``clustering(G)`` on the DiGraph restated (utils.py:56-60 call site; Fagiolo's directed
    clustering, ``_directed_triangles_and_degree_iter``): integer counts, one true division."""
    n = int(num_nodes)
    s, d = dedup_edges(edge_index, n)
    succ = [set() for _ in range(n)]
    pred = [set() for _ in range(n)]
    for u, v in zip(s.tolist(), d.tolist()):
        if u != v:
            succ[u].add(v)
            pred[v].add(u)
    out = np.zeros(n, dtype=np.float64)
    for i in range(n):
        P, S = pred[i], succ[i]
        t = 0
        for j in list(P) + list(S):
            t += len(P & pred[j]) + len(P & succ[j]) + len(S & pred[j]) + len(S & succ[j])
        dt = len(P) + len(S)
        db = len(P & S)
        out[i] = 0.0 if t == 0 else t / ((dt * (dt - 1) - 2 * db) * 2)
    return out


def clustering_coefficient_anchors(edge_index, num_nodes: int, k: int) -> list:
    return stable_top_k(clustering_scores(edge_index, num_nodes), k)
