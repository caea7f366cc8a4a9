"""Pin the sampler restatements to anchor lists produced by the reference."""
import numpy as np

from graphpope_b200 import synth
from oracle import samplers as s


def test_stochastic_matches_reference_stream(golden_small):
    np.random.seed(42)
    assert np.array_equal(s.stochastic(89250, 256), golden_small["samplers/stochastic_89250_256"])
    assert np.array_equal(synth.stochastic_anchors(89250, 256, 42)[:8],
                          [15795, 860, 76820, 54886, 6265, 82386, 37194, 87498])
    # np.random.choice(arange(N), K) consumes the same stream as randint(0, N, K)
    np.random.seed(42)
    assert np.array_equal(np.random.randint(0, 89250, 256), golden_small["samplers/stochastic_89250_256"])


def test_degree_centrality_matches_reference(golden_small):
    ei = golden_small["samplers/edge_index"]
    n = int(golden_small["samplers/n"])
    assert np.array_equal(s.degree_scores(ei, n), golden_small["samplers/degree"])
    for k in (1, 16, 64, 256):
        assert s.degree_centrality_anchors(ei, n, k) == golden_small[f"samplers/degree_centrality/{k}"].tolist()


def test_pagerank_matches_reference(golden_small):
    ei = golden_small["samplers/edge_index"]
    n = int(golden_small["samplers/n"])
    x, iters = s.pagerank_scores(ei, n)
    assert 1 <= iters <= 100
    want = golden_small["samplers/pagerank_scores"]
    assert np.abs(x - want).sum() < 1e-12
    assert np.array_equal(x, want)  # same operation order -> bit-equal float64
    for k in (1, 16, 64, 256):
        assert s.pagerank_anchors(ei, n, k) == golden_small[f"samplers/pagerank/{k}"].tolist()


def test_top_k_quirks():
    score = np.array([3, 1, 3, 2, 3])
    assert s.stable_top_k(score, 2) == [2, 4]          # ties -> larger ids, ascending score order
    assert s.stable_top_k(score, 0) == [1, 3, 0, 2, 4]  # list[-0:] is the whole list
    assert s.stable_top_k(score, 9) == [1, 3, 0, 2, 4]


def test_closeness_matches_reference(golden_small):
    ei = golden_small["samplers/edge_index"]
    n = int(golden_small["samplers/n"])
    x = s.closeness_scores(ei, n)
    assert np.array_equal(x, golden_small["samplers/closeness_scores"])  # same operation order -> bit-equal float64
    for k in (1, 16, 64, 256):
        assert s.closeness_centrality_anchors(ei, n, k) == golden_small[f"samplers/closeness_centrality/{k}"].tolist()


def test_clustering_matches_reference(golden_small):
    ei = golden_small["samplers/edge_index"]
    n = int(golden_small["samplers/n"])
    x = s.clustering_scores(ei, n)
    assert np.array_equal(x, golden_small["samplers/clustering_scores"])  # integer counts, one division
    for k in (1, 16, 64, 256):
        assert s.clustering_coefficient_anchors(ei, n, k) == golden_small[f"samplers/clustering_coefficient/{k}"].tolist()


def test_betweenness_matches_reference(golden_betweenness):
    g = golden_betweenness
    ei, n = g["edge_index"], int(g["n"])
    x = s.betweenness_scores(ei, n)
    assert np.array_equal(x, g["scores"])  # networkx's own summation order -> bit-equal float64
    for k in (1, 16, 64, 256):
        assert s.betweenness_centrality_anchors(ei, n, k) == g[f"anchors/{k}"].tolist()


def test_betweenness_device_summation_order_stays_within_a_few_ulp():
    """The device kernel sums the same terms level by level with neighbours in ascending id; restated on
    the CPU, that order agrees with networkx's queue order to ~1e-15 relative on an asymmetric digraph."""
    n = 120
    ei = synth.random_digraph(n, 420, seed=5)
    want = s.betweenness_scores(ei, n)
    got = s.betweenness_levelsync_scores(ei, n)
    assert np.array_equal(got == 0, want == 0)  # exact zeros (nodes on no shortest path) stay exact
    assert np.allclose(got, want, rtol=1e-13, atol=0)


def test_eigenvector_matches_reference(golden_eigenvector):
    import pytest
    g = golden_eigenvector
    ei, n = g["edge_index"], int(g["n"])
    x = s.eigenvector_scores(ei, n)
    # ARPACK starts from a random vector: the reference itself reproduces these scores only to ~1e-16
    assert np.allclose(x, g["scores"], rtol=0, atol=1e-12)
    for k in (1, 16, 64, 256):
        assert s.stable_top_k(x, k) == g[f"anchors/{k}"].tolist()
    with pytest.raises(ValueError):  # networkx >= 3.2 raises AmbiguousSolution on a disconnected graph
        s.eigenvector_scores(np.array([[0, 1], [1, 0]]), 3)


def _random_small_digraph(seed, n, m):
    rng = np.random.default_rng(seed)
    return rng.integers(0, n, size=(2, m)).astype(np.int64)


def test_betweenness_restatements_against_networkx_on_random_multigraphs():
    """Self-loops, parallel edges, isolated nodes, unreachable pairs: the queue-order restatement is bit-equal to
    networkx, the level-synchronous order (what the device kernel sums) within a few ulp with the same zeros."""
    import networkx as nx
    for seed in range(12):
        n = 5 + 3 * seed
        ei = _random_small_digraph(seed, n, 2 * n + seed)
        G = nx.DiGraph()
        G.add_nodes_from(range(n))
        G.add_edges_from(zip(ei[0].tolist(), ei[1].tolist()))
        want = nx.betweenness_centrality(G)
        want = np.asarray([want[i] for i in range(n)])
        assert np.array_equal(s.betweenness_scores(ei, n), want), seed
        lv = s.betweenness_levelsync_scores(ei, n)
        assert np.array_equal(lv == 0, want == 0), seed
        assert np.allclose(lv, want, rtol=1e-13, atol=0), seed


def test_eigenvector_restatement_against_networkx_on_strongly_connected_graphs():
    import networkx as nx
    for seed in range(6):
        n = 12 + 5 * seed
        ring = np.stack([np.arange(n), (np.arange(n) + 1) % n])  # a directed cycle keeps it strongly connected
        ei = np.concatenate([ring, _random_small_digraph(100 + seed, n, 3 * n)], axis=1)
        G = nx.DiGraph()
        G.add_nodes_from(range(n))
        G.add_edges_from(zip(ei[0].tolist(), ei[1].tolist()))
        want = nx.eigenvector_centrality_numpy(G)
        want = np.asarray([want[i] for i in range(n)])
        assert np.allclose(s.eigenvector_scores(ei, n), want, rtol=0, atol=1e-12), seed
