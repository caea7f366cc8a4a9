"""CPU restatement of the anchor samplers (reference utils.py:18-62).

TEST INFRASTRUCTURE — see oracle/__init__.py.

Only the samplers the device path re-implements are restated here
(all seven: stochastic, degree_centrality, pagerank, closeness_centrality,
clustering_coefficient, betweenness_centrality, eigenvector_centrality).

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


def betweenness_scores(edge_index, num_nodes: int) -> np.ndarray:
    """networkx ``betweenness_centrality(G)`` defaults restated (utils.py:32-36 call site): Brandes'
    algorithm per source exactly as ``_single_source_shortest_path_basic`` / ``_accumulate_basic`` /
    ``_rescale(normalized=True, directed=True, endpoints=False)`` do it, adjacency in ``to_networkx``
    insertion order (first appearance of each edge in ``edge_index``), so the float64 sums run in the
    same order and the scores are bit-equal to networkx's.  Pure-Python loops: small graphs only.
    """
    from collections import deque

    n = int(num_nodes)
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    adj = [dict() for _ in range(n)]  # insertion-ordered successor sets, as the DiGraph keeps them
    for u, v in zip(ei[0].tolist(), ei[1].tolist()):
        adj[u][v] = None
    adj = [list(a) for a in adj]
    bc = [0.0] * n
    for s in range(n):
        S = []
        P = [[] for _ in range(n)]
        sigma = [0.0] * n
        D = {}
        sigma[s] = 1.0
        D[s] = 0
        Q = deque([s])
        while Q:
            v = Q.popleft()
            S.append(v)
            Dv = D[v]
            sigmav = sigma[v]
            for w in adj[v]:
                if w not in D:
                    Q.append(w)
                    D[w] = Dv + 1
                if D[w] == Dv + 1:
                    sigma[w] += sigmav
                    P[w].append(v)
        delta = dict.fromkeys(S, 0)
        while S:
            w = S.pop()
            coeff = (1 + delta[w]) / sigma[w]
            for v in P[w]:
                delta[v] += sigma[v] * coeff
            if w != s:
                bc[w] += delta[w]
    N = n - 1
    if N >= 2:
        scale = 1 / (N * (N - 1))
        if scale != 1:
            bc = [b * scale for b in bc]
    return np.asarray(bc, dtype=np.float64)


def betweenness_levelsync_scores(edge_index, num_nodes: int) -> np.ndarray:
    """The SAME quantity in the order the device kernel (gp_betweenness.cu) sums it: level-synchronous
    pull sweeps, neighbours in ascending id, ``coeff = (1 + delta) / sigma`` stored per finalised node,
    32 sources per batch whose deltas meet in a butterfly sum.  Used by the CPU tests to show that this
    re-ordering stays within a few ulp of the networkx order (the device itself is checked on the GPU).
    """
    n = int(num_nodes)
    s_, d_ = dedup_edges(edge_index, n)
    succ = [[] for _ in range(n)]
    pred = [[] for _ in range(n)]
    for u, v in sorted(zip(s_.tolist(), d_.tolist())):
        succ[u].append(v)
    for u, v in sorted(zip(s_.tolist(), d_.tolist()), key=lambda e: (e[1], e[0])):
        pred[v].append(u)
    bc = np.zeros(n, dtype=np.float64)
    for b0 in range(0, n, 32):
        lanes = list(range(b0, min(n, b0 + 32)))
        per_level = {}  # (level, row) -> [32 lane deltas]
        maxl_all = 0
        state = []
        for s in lanes:
            dist = [-1] * n
            sigma = [0.0] * n
            dist[s] = 0
            sigma[s] = 1.0
            lvl = 0
            while True:
                new = []
                for w in range(n):
                    if dist[w] != -1:
                        continue
                    acc = 0.0
                    for v in pred[w]:
                        if dist[v] == lvl:
                            acc = acc + sigma[v]
                    if acc != 0.0:
                        new.append((w, acc))
                for w, acc in new:
                    dist[w] = lvl + 1
                    sigma[w] = acc
                if not new:
                    break
                lvl += 1
            state.append((dist, sigma))
            maxl_all = max(maxl_all, lvl)
        coeffs = [[0.0] * n for _ in lanes]
        for l in range(maxl_all - 1, 0, -1):
            for v in range(n):
                vals = [0.0] * 32
                hit = False
                for li, (dist, sigma) in enumerate(state):
                    if dist[v] != l:
                        continue
                    hit = True
                    acc = 0.0
                    for w in succ[v]:
                        if dist[w] == l + 1:
                            cw = 1.0 / sigma[w] if l + 1 == maxl_all else coeffs[li][w]
                            acc = acc + sigma[v] * cw
                    coeffs[li][v] = (1.0 + acc) / sigma[v]
                    vals[li] = acc
                if hit:
                    m = 16
                    while m >= 1:  # butterfly: every lane ends with the same total
                        vals = [vals[i] + vals[i ^ m] for i in range(32)]
                        m >>= 1
                    bc[v] = bc[v] + vals[0]
    if n - 1 >= 2:
        scale = 1.0 / (float(n - 1) * float(n - 2))
        if scale != 1.0:
            bc = bc * scale
    return bc


def betweenness_centrality_anchors(edge_index, num_nodes: int, k: int) -> list:
    return stable_top_k(betweenness_scores(edge_index, num_nodes), k)


def eigenvector_scores(edge_index, num_nodes: int) -> np.ndarray:
    """networkx ``eigenvector_centrality_numpy(G)`` defaults restated (utils.py:44-48 call site; networkx
    3.6.1): refuse graphs that are not strongly connected, hand ``A^T`` to ARPACK (``eigs(k=1, which='LR',
    maxiter=50, tol=0)``), take the real part, scale to unit L2 norm with a positive sum.  ARPACK starts from
    a random vector, so even two runs of the reference differ in the last bits (~1e-16): parity for this
    sampler is a tolerance (tests: 1e-12 against the frozen scores), not bit equality.
    """
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    from scipy.sparse.csgraph import connected_components

    n = int(num_nodes)
    if n == 0:
        raise ValueError("cannot compute centrality for the null graph")
    s_, d_ = dedup_edges(edge_index, n)
    M = sp.csr_array((np.ones(s_.size), (s_, d_)), shape=(n, n))
    ncomp, _ = connected_components(M, directed=True, connection="strong")
    if ncomp != 1:
        raise ValueError("`eigenvector_centrality_numpy` does not give consistent results for disconnected graphs")
    _, vec = sla.eigs(M.T, k=1, which="LR", maxiter=50, tol=0)
    largest = vec.flatten().real
    norm = np.sign(largest.sum()) * np.linalg.norm(largest)
    return largest / norm


def eigenvector_centrality_anchors(edge_index, num_nodes: int, k: int) -> list:
    return stable_top_k(eigenvector_scores(edge_index, num_nodes), k)
