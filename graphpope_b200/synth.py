"""Synthetic inputs shaped like the BASELINE.json configs (there is no network).

Every graph is an undirected simple graph stored the way PyG ships Flickr /
PubMed / ogbn-products: a *symmetric directed* ``edge_index`` int64 ``[2, E]``
with no self-loops and no duplicate columns, column order shuffled
(SURVEY.md §8d).  Degrees follow a Chung-Lu model with Pareto endpoint weights;
pairs are resampled until the exact directed edge count is reached.

Only numpy is used so the same generator feeds the CPU oracle, the tests and
``bench.py``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class GraphShape:
    name: str
    num_nodes: int
    num_directed_edges: int  # E of the symmetric directed edge_index
    pareto_alpha: float
    num_features: int
    seed: int


# C1, C2/C3/C4 and C5.
PUBMED_SHAPE = GraphShape("pubmed-shape", 19_717, 88_648, 2.3, 500, 1)
FLICKR_SHAPE = GraphShape("flickr-shape", 89_250, 899_756, 1.9, 500, 2)
PRODUCTS_SHAPE = GraphShape("products-shape", 2_449_029, 123_718_280, 2.1, 100, 5)

SHAPES = {s.name: s for s in (PUBMED_SHAPE, FLICKR_SHAPE, PRODUCTS_SHAPE)}


def chung_lu_symmetric(num_nodes: int, num_directed_edges: int, alpha: float,
                       seed: int) -> np.ndarray:
    """Symmetric directed edge_index ``int64[2, E]`` with exactly E columns."""
    if num_directed_edges % 2:
        raise ValueError("a symmetric edge_index has an even number of columns")
    n = int(num_nodes)
    target = num_directed_edges // 2
    if target > n * (n - 1) // 2:
        raise ValueError("more edges requested than a simple graph can hold")
    rng = np.random.default_rng(seed)
    w = rng.pareto(alpha, n) + 1.0
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    have = np.empty(0, dtype=np.int64)  # encoded lo * n + hi, lo < hi, unique
    while have.size < target:
        need = target - have.size
        m = int(need * 1.25) + 16
        a = np.searchsorted(cdf, rng.random(m), side="right").astype(np.int64)
        b = np.searchsorted(cdf, rng.random(m), side="right").astype(np.int64)
        np.minimum(a, n - 1, out=a)
        np.minimum(b, n - 1, out=b)
        keep = a != b
        lo = np.minimum(a[keep], b[keep])
        hi = np.maximum(a[keep], b[keep])
        have = np.unique(np.concatenate([have, lo * n + hi]))
        if have.size > target:
            # np.unique sorted the codes; drop a random subset, not the tail
            have = have[np.sort(rng.choice(have.size, target, replace=False))]
    lo = have // n
    hi = have % n
    src = np.concatenate([lo, hi])
    dst = np.concatenate([hi, lo])
    perm = rng.permutation(src.size)
    return np.stack([src[perm], dst[perm]]).astype(np.int64)


def make_graph(shape: GraphShape | str) -> np.ndarray:
    if isinstance(shape, str):
        shape = SHAPES[shape]
    return chung_lu_symmetric(shape.num_nodes, shape.num_directed_edges,
                              shape.pareto_alpha, shape.seed)


def random_digraph(num_nodes: int, num_edges: int, seed: int,
                   self_loops: bool = True, duplicates: bool = True) -> np.ndarray:
    """Small *asymmetric* multigraph for parity tests (direction, dedup, loops)."""
    rng = np.random.default_rng(seed)
    if num_nodes == 0 or num_edges == 0:
        return np.zeros((2, 0), dtype=np.int64)
    src = rng.integers(0, num_nodes, num_edges)
    dst = rng.integers(0, num_nodes, num_edges)
    if not self_loops:
        keep = src != dst
        src, dst = src[keep], dst[keep]
    if duplicates and src.size:
        k = max(1, src.size // 8)
        idx = rng.integers(0, src.size, k)
        src = np.concatenate([src, src[idx]])
        dst = np.concatenate([dst, dst[idx]])
    return np.stack([src, dst]).astype(np.int64)


def stochastic_anchors(num_nodes: int, k: int, seed: int = 42) -> np.ndarray:
    """Anchors exactly as utils.py:22-24 draws them after ``seed_everything``.

    ``np.random.choice(np.arange(N), K)`` on the legacy global RNG, with
    replacement (duplicates are kept and become duplicate columns).
    """
    st = np.random.get_state()
    try:
        np.random.seed(seed)
        return np.random.choice(np.arange(num_nodes), k).astype(np.int64)
    finally:
        np.random.set_state(st)


def node2vec_table(num_nodes: int, dim: int = 128, seed: int = 3) -> np.ndarray:
    """Stand-in for ``data/<dataset>_node2vec.pt``.

    The reference's generator never trains the model
    (generate_node2vec_embedding.py:23-28), so its file is the N(0,1) initial
    embedding table; a Gaussian table is therefore a faithful synthetic input.
    """
    return np.random.default_rng(seed).standard_normal((num_nodes, dim)).astype(np.float32)


def chung_lu_symmetric_torch(num_nodes: int, num_directed_edges: int, alpha: float, seed: int, device="cuda"):
    """The same degree model as :func:`chung_lu_symmetric`, drawn with torch on ``device`` (a different random
    stream, so a different graph of the same shape): symmetric directed ``edge_index`` int64 ``[2, E]`` with exactly
    E columns, no self-loops, no duplicates, columns shuffled.  The ogbn-products-shaped graph (2.45 M nodes,
    123.7 M directed edges) takes ~110 s with numpy on the host and about a second here.  Bench / test input only."""
    import torch

    if num_directed_edges % 2:
        raise ValueError("a symmetric edge_index has an even number of columns")
    n, target = int(num_nodes), num_directed_edges // 2
    g = torch.Generator(device=device).manual_seed(int(seed))
    u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
    w = (1.0 - u).pow(-1.0 / alpha)  # Pareto(alpha) + 1, as numpy's rng.pareto(alpha) + 1
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    have = torch.empty(0, dtype=torch.int64, device=device)
    while have.numel() < target:
        need = target - have.numel()
        m = int(need * 1.25) + 16
        a = torch.searchsorted(cdf, torch.rand(m, generator=g, device=device, dtype=torch.float64), right=True).clamp_(max=n - 1)
        b = torch.searchsorted(cdf, torch.rand(m, generator=g, device=device, dtype=torch.float64), right=True).clamp_(max=n - 1)
        keep = a != b
        lo, hi = torch.minimum(a[keep], b[keep]), torch.maximum(a[keep], b[keep])
        have = torch.unique(torch.cat([have, lo * n + hi]))
        if have.numel() > target:
            sel = torch.randperm(have.numel(), generator=g, device=device)[:target]
            have = have[sel.sort().values]
    lo, hi = have // n, have % n
    src, dst = torch.cat([lo, hi]), torch.cat([hi, lo])
    perm = torch.randperm(src.numel(), generator=g, device=device)
    return torch.stack([src[perm], dst[perm]])
