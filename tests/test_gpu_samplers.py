"""Device degree / PageRank samplers and stable top-k against anchor lists produced by the reference."""
import numpy as np
import pytest
import torch

from graphpope_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from graphpope_b200 import device
    return device


class Data:
    def __init__(self, ei, n):
        self.edge_index, self.num_nodes = torch.as_tensor(ei), n


def test_degree_scores_and_anchor_lists(dev, golden_small):
    from graphpope_b200 import utils
    ei, n = golden_small["samplers/edge_index"], int(golden_small["samplers/n"])
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    assert np.array_equal(csr.degree().cpu().numpy(), golden_small["samplers/degree"])
    for k in (1, 16, 64, 256):
        assert utils.sample_anchor_nodes(Data(ei, n), k, "degree_centrality") == \
            golden_small[f"samplers/degree_centrality/{k}"].tolist()


def test_pagerank_scores_bit_equal_and_anchor_lists(dev, golden_small):
    from graphpope_b200 import utils
    ei, n = golden_small["samplers/edge_index"], int(golden_small["samplers/n"])
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    x, iters = csr.pagerank()
    want = golden_small["samplers/pagerank_scores"]
    got = x.cpu().numpy()
    assert 1 <= iters <= 100
    assert np.abs(got - want).sum() < 1e-12
    assert np.array_equal(got, want)  # same operation order as scipy -> identical float64
    for k in (1, 16, 64, 256):
        assert utils.sample_anchor_nodes(Data(ei, n), k, "pagerank") == golden_small[f"samplers/pagerank/{k}"].tolist()


def test_pagerank_matches_oracle_on_directed_graph_with_dangling_nodes(dev):
    from oracle import samplers as s
    n = 3000
    ei = synth.random_digraph(n, 7000, seed=11)  # asymmetric, many nodes without out-edges
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    x, iters = csr.pagerank()
    want, it_want = s.pagerank_scores(ei, n)
    assert iters == it_want
    assert np.array_equal(x.cpu().numpy(), want)
    assert dev.topk_stable(x, 50).cpu().tolist() == s.stable_top_k(want, 50)


def test_topk_stable_quirks(dev):
    score = torch.tensor([3, 1, 3, 2, 3], dtype=torch.int32).cuda()
    assert dev.topk_stable(score, 2).cpu().tolist() == [2, 4]           # ties -> larger ids, ascending score
    assert dev.topk_stable(score, 0).cpu().tolist() == [1, 3, 0, 2, 4]  # list[-0:] is everything
    assert dev.topk_stable(score, 9).cpu().tolist() == [1, 3, 0, 2, 4]
    f = torch.tensor([0.5, -1.0, 0.5, 0.0, 2.0], dtype=torch.float64).cuda()
    assert dev.topk_stable(f, 3).cpu().tolist() == [0, 2, 4]
    big = torch.randint(0, 50, (100000,), dtype=torch.int32, device="cuda")
    want = np.argsort(big.cpu().numpy(), kind="stable")[-1000:]
    assert np.array_equal(dev.topk_stable(big, 1000).cpu().numpy(), want)


def test_flickr_shape_degree_and_pagerank_1024_anchors(dev):
    """BASELINE.json configs[3] samplers on the Flickr-shape graph against the CPU oracle."""
    from oracle import samplers as s
    shape = synth.FLICKR_SHAPE
    ei = synth.make_graph(shape)
    csr = dev.DeviceCsr(shape.num_nodes, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    deg = csr.degree()
    assert np.array_equal(deg.cpu().numpy(), s.degree_scores(ei, shape.num_nodes))
    assert dev.topk_stable(deg, 1024).cpu().tolist() == s.degree_centrality_anchors(ei, shape.num_nodes, 1024)
    x, _ = csr.pagerank()
    assert abs(float(x.sum()) - 1.0) < 1e-9


def test_closeness_scores_bit_equal_and_anchor_lists(dev, golden_small):
    """nx.closeness_centrality (utils.py:50-54) from the MS-BFS with every node as an anchor."""
    from graphpope_b200 import utils
    ei, n = golden_small["samplers/edge_index"], int(golden_small["samplers/n"])
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    got = csr.closeness().cpu().numpy()
    assert np.array_equal(got, golden_small["samplers/closeness_scores"])  # float64, bit for bit
    for k in (1, 16, 64, 256):
        assert utils.sample_anchor_nodes(Data(ei, n), k, "closeness_centrality") == \
            golden_small[f"samplers/closeness_centrality/{k}"].tolist()


@pytest.mark.parametrize("case", ["directed-sparse", "path", "multi-chunk"])
def test_closeness_matches_oracle_on_hard_graphs(dev, case):
    """Unreachable pairs and dangling nodes (asymmetric digraph), hop counts beyond 15 (a path: the
    deep bit planes and the wide adder), and more nodes than one 4096-anchor pass."""
    from oracle import samplers as s
    if case == "directed-sparse":
        n = 700
        ei = synth.random_digraph(n, 1500, seed=31)
    elif case == "path":
        n = 300
        a = np.arange(n - 1)
        ei = np.stack([np.concatenate([a, a + 1]), np.concatenate([a + 1, a])]).astype(np.int64)
    else:
        n = 9000
        ei = synth.chung_lu_symmetric(n, 40000, 2.2, seed=33)
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    got = csr.closeness().cpu().numpy()
    assert np.array_equal(got, s.closeness_scores(ei, n))


def test_clustering_scores_bit_equal_and_anchor_lists(dev, golden_small):
    """nx.clustering on the DiGraph (utils.py:56-60) from shared-memory node bitmaps."""
    from graphpope_b200 import utils
    ei, n = golden_small["samplers/edge_index"], int(golden_small["samplers/n"])
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    got = csr.clustering().cpu().numpy()
    assert np.array_equal(got, golden_small["samplers/clustering_scores"])
    for k in (1, 16, 64, 256):
        assert utils.sample_anchor_nodes(Data(ei, n), k, "clustering_coefficient") == \
            golden_small[f"samplers/clustering_coefficient/{k}"].tolist()


def test_clustering_matches_oracle_on_a_dense_directed_multigraph(dev):
    """Self-loops, reciprocal edges, duplicates and a hub: every term of the directed formula."""
    from oracle import samplers as s
    n = 400
    rng = np.random.default_rng(8)
    ei = synth.random_digraph(n, 6000, seed=41)
    hub = np.stack([np.zeros(300, dtype=np.int64), rng.choice(n, 300, replace=False)])
    ei = np.concatenate([ei, hub, hub[::-1][:, :150]], axis=1)
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    assert np.array_equal(csr.clustering().cpu().numpy(), s.clustering_scores(ei, n))


def _assert_same_ranking(got, want, k, rtol=1e-10):
    """Top-k lists agree wherever the reference's own scores separate the candidates by more than rtol."""
    from oracle import samplers as s
    a, b = s.stable_top_k(got, k), s.stable_top_k(want, k)
    if a == b:
        return
    cut = want[b[0]]  # smallest selected reference score
    for u, v in zip(a, b):
        if u != v:
            assert abs(want[u] - want[v]) <= rtol * max(abs(want[u]), abs(want[v]), 1e-300) or \
                abs(want[u] - cut) <= rtol * abs(cut), (u, v, want[u], want[v])


def _assert_anchor_lists(method, ei, n, golden, want_scores, rtol):
    """utils.sample_anchor_nodes against the lists the unmodified reference returned, for every K frozen in the
    fixture; positions may differ only where the reference's own scores tie within the sampler's tolerance."""
    from graphpope_b200 import utils
    for k in (1, 16, 64, 256):
        got = utils.sample_anchor_nodes(Data(ei, n), k, method)
        want = golden[f"anchors/{k}"].tolist()
        assert len(got) == len(want)
        for u, v in zip(got, want):
            if u != v:
                assert abs(want_scores[u] - want_scores[v]) <= rtol * max(abs(want_scores[u]), abs(want_scores[v]), 1e-300), \
                    (method, k, u, v)


def test_betweenness_scores_and_anchor_lists(dev, golden_betweenness):
    """nx.betweenness_centrality (utils.py:32-36): Brandes with 32 sources per warp-wide batch."""
    from graphpope_b200 import utils
    g = golden_betweenness
    ei, n = g["edge_index"], int(g["n"])
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    got = csr.betweenness().cpu().numpy()
    want = g["scores"]
    assert np.array_equal(got == 0, want == 0)
    assert np.allclose(got, want, rtol=1e-10, atol=0)  # tolerance of this sampler: summation order differs
    again = csr.betweenness().cpu().numpy()
    assert np.array_equal(got, again)  # fixed summation order: run-to-run bit-equal
    for k in (1, 16, 64, 256):
        _assert_same_ranking(got, want, k)
    assert utils.sample_anchor_nodes(Data(ei, n), 16, "betweenness_centrality") == g["anchors/16"].tolist()
    _assert_anchor_lists("betweenness_centrality", ei, n, g, want, 1e-10)


@pytest.mark.parametrize("spl", [1, 2, 4])
@pytest.mark.parametrize("case", ["directed-sparse", "path", "hub", "tiny"])
def test_betweenness_matches_oracle_on_hard_graphs(dev, case, spl, monkeypatch):
    """Unreachable pairs and dangling nodes; a 300-node path (hop counts beyond 253: the 16-bit restart);
    a star-of-stars whose centre row is cut into 64-edge chunks; graphs of 1..3 nodes (no rescale) — each
    with 1, 2 and 4 sources per lane (32 / 64 / 128 sources per batch)."""
    from oracle import samplers as s
    monkeypatch.setenv("GP_BC_SOURCES", str(spl))
    if case == "directed-sparse":
        n = 700
        ei = synth.random_digraph(n, 1500, seed=31)
    elif case == "path":
        n = 300
        a = np.arange(n - 1)
        ei = np.stack([np.concatenate([a, a + 1]), np.concatenate([a + 1, a])]).astype(np.int64)
    elif case == "hub":
        n = 500
        spokes = np.arange(1, 300)
        leaves = np.arange(300, 500)
        src = np.concatenate([np.zeros_like(spokes), spokes, spokes[:200], leaves])
        dst = np.concatenate([spokes, np.zeros_like(spokes), leaves, spokes[:200]])
        extra = synth.random_digraph(n, 300, seed=7)
        ei = np.concatenate([np.stack([src, dst]).astype(np.int64), extra], axis=1)
    else:
        for n, edges in ((1, [(0, 0)]), (2, [(0, 1), (1, 0)]), (3, [(0, 1), (1, 2)])):
            ei = np.asarray(edges, dtype=np.int64).T.copy()
            csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
            assert np.array_equal(csr.betweenness().cpu().numpy(), s.betweenness_scores(ei, n))
        return
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    got = csr.betweenness().cpu().numpy()
    want = s.betweenness_scores(ei, n)
    assert np.array_equal(got == 0, want == 0)
    assert np.allclose(got, want, rtol=1e-10, atol=0)


def test_eigenvector_scores_and_anchor_lists(dev, golden_eigenvector):
    """nx.eigenvector_centrality_numpy (utils.py:44-48) by float64 power iteration on A^T + I."""
    import networkx as nx

    from graphpope_b200 import utils
    g = golden_eigenvector
    ei, n = g["edge_index"], int(g["n"])
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    got, iters = csr.eigenvector()
    got = got.cpu().numpy()
    assert 1 <= iters < 20000
    assert abs(np.linalg.norm(got) - 1.0) < 1e-14 and (got > 0).all()
    assert np.allclose(got, g["scores"], rtol=1e-9, atol=1e-12)  # tolerance of this sampler (ARPACK vs power iteration)
    for k in (1, 16, 64, 256):
        _assert_same_ranking(got, g["scores"], k, rtol=1e-9)
    assert utils.sample_anchor_nodes(Data(ei, n), 16, "eigenvector_centrality") == g["anchors/16"].tolist()
    _assert_anchor_lists("eigenvector_centrality", ei, n, g, g["scores"], 1e-9)
    # a directed cycle (period 5: every eigenvalue has modulus 1, the shift makes the Perron root dominant)
    cyc = np.array([[0, 1, 2, 3, 4], [1, 2, 3, 4, 0]])
    x, _ = dev.DeviceCsr(5, 5).build(torch.as_tensor(cyc).cuda()).eigenvector()
    assert np.allclose(x.cpu().numpy(), 1 / np.sqrt(5), rtol=0, atol=1e-12)
    # not strongly connected: the reference (networkx >= 3.2) raises AmbiguousSolution, and so does the mirror
    with pytest.raises(nx.AmbiguousSolution):
        utils.sample_anchor_nodes(Data(np.array([[0, 1, 2], [1, 0, 1]]), 4), 2, "eigenvector_centrality")
