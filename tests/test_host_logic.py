"""Host-side logic of the reference-facing mirror (no GPU): signatures, errors, memo, samplers that stay on the host."""
import inspect

import numpy as np
import pytest
import torch

from graphpope_b200 import utils


class Data:
    def __init__(self, edge_index, num_nodes, x=None):
        self.edge_index = torch.as_tensor(edge_index)
        self.num_nodes = num_nodes
        self.x = x


def test_signatures_match_the_reference():
    # utils.py:18,64,83,92,116,129,137,149,182
    want = {
        "sample_anchor_nodes": ["data", "num_anchor_nodes", "sampling_method"],
        "shortest_path_length": ["G", "anchor_nodes", "partition_length"],
        "merge_dicts": ["dicts"],
        "all_pairs_shortest_path_length_parallel": ["G", "anchor_nodes", "num_workers"],
        "get_geodesic_distance_vector": ["data", "num_workers"],
        "concat_into_features": ["embedding_matrix", "data"],
        "attach_distance_embedding": ["data", "dataset", "num_anchor_nodes", "sampling_method",
                                      "distance_function", "num_workers"],
        "attach_node2vec": ["data", "dataset", "num_anchor_nodes", "sampling_method", "distance_function",
                            "num_workers"],
        "Graphpope": ["data", "dataset", "embedding_space", "sampling_method", "num_anchor_nodes",
                      "distance_function", "num_workers"],
    }
    for name, params in want.items():
        assert list(inspect.signature(getattr(utils, name)).parameters) == params, name
    sig = inspect.signature(utils.Graphpope)
    assert sig.parameters["distance_function"].default is None and sig.parameters["num_workers"].default == 4


def test_stochastic_sampler_is_the_reference_stream(golden_small):
    np.random.seed(42)
    got = utils.sample_anchor_nodes(Data(np.zeros((2, 0), np.int64), 89250), 256, "stochastic")
    assert isinstance(got, np.ndarray)
    assert np.array_equal(got, golden_small["samplers/stochastic_89250_256"])


def test_unknown_sampler_raises_like_the_reference():
    with pytest.raises(UnboundLocalError):
        utils.sample_anchor_nodes(Data(np.zeros((2, 0), np.int64), 5), 2, "kmeans")


def test_host_centralities_use_networkx_top_k_rule(monkeypatch):
    import networkx as nx
    ei = np.array([[0, 1, 1, 2, 2, 3, 3, 4, 1, 3], [1, 0, 2, 1, 3, 2, 4, 3, 3, 1]])
    data = Data(ei, 6)
    if not torch.cuda.is_available():  # betweenness runs on the device by default and must fail loudly without one
        with pytest.raises(RuntimeError):
            utils.sample_anchor_nodes(data, 3, "betweenness_centrality")
    monkeypatch.setenv("GRAPHPOPE_BETWEENNESS", "networkx")  # the reference's own call, kept as an option
    got = utils.sample_anchor_nodes(data, 3, "betweenness_centrality")
    G = nx.DiGraph(); G.add_nodes_from(range(6)); G.add_edges_from(zip(*ei.tolist()))
    score = nx.betweenness_centrality(G)
    want = [k for k, _ in sorted(score.items(), key=lambda kv: kv[1])][-3:]
    assert got == want
    if not torch.cuda.is_available():  # closeness / clustering run on the device and must fail loudly without one
        for method in ("closeness_centrality", "clustering_coefficient"):
            with pytest.raises(RuntimeError):
                utils.sample_anchor_nodes(data, 3, method)


def test_merge_dicts_and_concat():
    assert utils.merge_dicts([{0: [1]}, {1: [2], 0: [3]}]) == {0: [3], 1: [2]}
    assert list(utils.merge_dicts([{2: 0}, {1: 0}])) == [2, 1]
    d = Data(np.zeros((2, 0), np.int64), 2, torch.tensor([[1.0, 2.0], [3.0, 4.0]]))
    out = utils.concat_into_features(np.array([[0.5], [0.25]], dtype=np.float32), d)
    assert out.tolist() == [[1.0, 2.0, 0.5], [3.0, 4.0, 0.25]]


def test_graphpope_dispatch_memo_and_errors(monkeypatch):
    utils.clear_cache()
    calls = []

    def fake(data, dataset, k, method, fn, num_workers):
        calls.append((dataset, k, method, fn, num_workers))
        return torch.full((2, 2), float(len(calls)))

    monkeypatch.setattr(utils, "attach_distance_embedding", fake)
    monkeypatch.setattr(utils, "attach_node2vec", fake)
    with pytest.raises(KeyError):
        utils.Graphpope(None, "flickr", "baseline", "stochastic", 2)
    a = utils.Graphpope(None, "flickr", "geodesic", "stochastic", 2, None, num_workers=6)
    b = utils.Graphpope(None, "pubmed", "node2vec", "kmeans", 99, "euclidean")
    assert a is b and calls == [("flickr", 2, "stochastic", None, 6)]  # memo ignores arguments (utils.py:202-208)
    utils.clear_cache()
    c = utils.Graphpope(None, "pubmed", "node2vec", "kmeans", 99, "euclidean")
    assert c is not a and calls[-1] == ("pubmed", 99, "kmeans", "euclidean", 4)
    utils.clear_cache()


def test_node2vec_unknown_distance_function_is_a_keyerror(tmp_path, monkeypatch):
    monkeypatch.setenv("GRAPHPOPE_DATA_DIR", str(tmp_path))
    torch.save(torch.zeros(4, 8), tmp_path / "toy_node2vec.pt")
    with pytest.raises(KeyError):
        utils.attach_node2vec(Data(np.zeros((2, 0), np.int64), 4, torch.zeros(4, 1)), "toy", 2, "stochastic", "None", 2)
    with pytest.raises(FileNotFoundError):
        utils.attach_node2vec(Data(np.zeros((2, 0), np.int64), 4, torch.zeros(4, 1)), "absent", 2, "stochastic",
                              "euclidean", 2)
    assert utils.node2vec_path("flickr").endswith("flickr_node2vec.pt")


def test_device_entry_points_fail_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = Data(np.array([[0], [1]]), 2, torch.zeros(2, 1))
    d.anchor_nodes = [0]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        utils.get_geodesic_distance_vector(d, 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        utils.sample_anchor_nodes(d, 1, "degree_centrality")


def test_host_concat_matches_torch_cat():
    """gp_host_concat (concat_into_features, utils.py:129-135, for host buffers) is pure host code: the
    worker pool + streaming-store row copy must equal torch.cat for aligned and odd shapes."""
    import ctypes

    import numpy as np
    import torch

    from graphpope_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n, f, k in ((1, 1, 1), (7, 3, 5), (1000, 500, 256), (40000, 37, 19), (3000, 500, 256)):
        x = torch.as_tensor(rng.standard_normal((n, f)).astype(np.float32))
        b = torch.as_tensor(rng.standard_normal((n, k)).astype(np.float32))
        out = torch.full((n, f + k), float("nan"))
        rc = lib.gp_host_concat(ctypes.c_void_p(x.data_ptr()), f, ctypes.c_void_p(b.data_ptr()), k, n,
                                ctypes.c_void_p(out.data_ptr()), f + k)
        assert rc == 0
        assert torch.equal(out, torch.cat((x, b), 1))
        # destination rows that start off a 16-byte boundary (a column slice of a wider buffer)
        wide = torch.full((n, f + k + 3), float("nan"))
        view = wide[:, 1:1 + f + k]
        rc = lib.gp_host_concat(ctypes.c_void_p(x.data_ptr()), f, ctypes.c_void_p(b.data_ptr()), k, n,
                                ctypes.c_void_p(view.data_ptr()), f + k + 3)
        assert rc == 0
        assert torch.equal(view, torch.cat((x, b), 1))
        assert torch.isnan(wide[:, 0]).all() and torch.isnan(wide[:, -2:]).all()


def _concat_in_child(q):
    import ctypes

    import numpy as np  # (no torch ops here: torch's own thread pool does not survive fork() either)

    from graphpope_b200 import _lib
    lib = _lib.load()
    n, f = 6000, 400  # 2.4 M elements: above the threshold that hands the copy to the worker pool
    x = np.arange(n * f, dtype=np.float32).reshape(n, f)
    out = np.zeros((n, f + 4), dtype=np.float32)
    rc = lib.gp_host_concat(ctypes.c_void_p(x.ctypes.data), f, None, 0, n, ctypes.c_void_p(out.ctypes.data), f + 4)
    q.put(rc == 0 and bool(np.array_equal(out[:, :f], x)))


def test_host_copy_pool_survives_fork():
    """The worker threads of the host row-copy pool do not exist in a forked child (DataLoader workers,
    multiprocessing): the child must copy the rows itself instead of waiting for them forever."""
    import ctypes
    import multiprocessing as mp

    import torch

    from graphpope_b200 import _lib
    lib = _lib.load()
    x = torch.ones(6000, 400)
    out = torch.zeros(6000, 404)
    assert lib.gp_host_concat(ctypes.c_void_p(x.data_ptr()), 400, None, 0, 6000, ctypes.c_void_p(out.data_ptr()), 404) == 0
    ctx = mp.get_context("fork")  # the pool now exists in this process
    q = ctx.Queue()
    p = ctx.Process(target=_concat_in_child, args=(q,))
    p.start()
    p.join(60)
    alive = p.is_alive()
    if alive:
        p.kill()
    assert not alive, "gp_host_concat deadlocked in a forked child"
    assert q.get(timeout=5) is True


def test_block_cache_key_and_hit_path(tmp_path, monkeypatch):
    """GRAPHPOPE_CACHE_DIR: the [N, K] block is stored once and served from disk after;
    the key depends on the graph, the anchors and the options."""
    ei = np.array([[0, 1, 2], [1, 2, 0]])
    k1 = utils.block_cache_key(ei, 3, [0, 2])
    assert k1 == utils.block_cache_key(torch.as_tensor(ei), 3, np.array([0, 2]))
    assert k1 != utils.block_cache_key(ei, 3, [2, 0])
    assert k1 != utils.block_cache_key(ei, 4, [0, 2])
    assert k1 != utils.block_cache_key(ei[:, ::-1].copy(), 3, [0, 2])
    assert k1 != utils.block_cache_key(ei, 3, [0, 2], symmetrize=True)

    calls = []

    def fake_embed(edge_index, n, anchors, x, sym):
        calls.append(1)
        return torch.arange(n * len(anchors), dtype=torch.float32).view(n, len(anchors)), None, {}

    monkeypatch.setenv("GRAPHPOPE_CACHE_DIR", str(tmp_path))
    monkeypatch.setattr(utils._dev, "geodesic_embed_host", fake_embed)
    d = Data(ei, 3, torch.ones(3, 2))
    d.anchor_nodes = [0, 2]
    a = utils.get_geodesic_distance_vector(d, 4)
    b = utils.get_geodesic_distance_vector(d, 4)
    assert len(calls) == 1 and torch.equal(a, b) and len(list(tmp_path.iterdir())) == 1
    d.anchor_nodes = [1, 2]
    utils.get_geodesic_distance_vector(d, 4)
    assert len(calls) == 2 and len(list(tmp_path.iterdir())) == 2
    # the concat route goes through the cache too
    monkeypatch.setattr(utils, "sample_anchor_nodes", lambda **kw: [0, 2])
    out = utils.attach_distance_embedding(d, "toy", 2, "stochastic", None, 4)
    assert len(calls) == 2 and out.shape == (3, 4) and torch.equal(out[:, 2:], a)


def test_shared_matrix_that_does_not_fit_is_an_error_not_a_crash():
    """SharedHostMatrix reserves its pages when it creates the file: asking for more than /dev/shm holds raises."""
    import os
    import shutil

    from graphpope_b200 import distributed as gpd

    total = shutil.disk_usage("/dev/shm").total
    rows = total // (4 * 1024) + (1 << 20)  # 1024 columns: ~4 GiB more than the whole file system (refused up front)
    before = set(os.listdir("/dev/shm"))
    with pytest.raises(RuntimeError, match="node-shared matrix"):
        gpd.SharedHostMatrix(rows, 1024, register=False)
    assert set(os.listdir("/dev/shm")) == before  # nothing left behind
    m = gpd.SharedHostMatrix(5, 7, register=False)  # single process, no process group
    m.tensor.fill_(2.0)
    assert m.tensor.shape == (5, 7) and float(m.tensor.sum()) == 70.0 and not os.path.exists(m.path)
    m.close()


@pytest.mark.parametrize("workers", [1, 2, 3])
def test_pool_helper_shims_host_logic(golden_small, monkeypatch, workers):
    """The reference-shaped halves of the pool helpers (utils.py:64-114) — node slices in float arithmetic, ordered
    merge, Python ``float`` / literal ``int 0`` entries — against what the unmodified reference returned
    (reference_shims.npz).  The hop matrix comes from the oracle here (no GPU); tests/test_gpu_shims.py runs the same
    checks with the device sweep underneath."""
    import os

    import networkx as nx

    from conftest import GOLDEN_DIR, micro_names
    from oracle import geodesic as og

    shims = np.load(os.path.join(GOLDEN_DIR, "reference_shims.npz"))

    def oracle_hops(G, anchor_nodes):
        ei, n = utils._graph_to_edge_index(G)
        return og.t2_bfs_csr_hops(ei.numpy(), n, [int(a) for a in anchor_nodes])

    monkeypatch.setattr(utils, "_hops_of_graph", oracle_hops)

    def check(got, keys, rows, is_int):
        assert list(got.keys()) == keys.tolist()
        for k, want_row, int_row in zip(keys.tolist(), rows, is_int):
            assert [type(v) for v in got[k]] == [int if i else float for i in int_row]
            assert got[k] == want_row.tolist()

    for name in micro_names(golden_small):
        n = int(golden_small[f"micro/{name}/n"])
        ei = golden_small[f"micro/{name}/edges"]
        G = nx.DiGraph()
        G.add_nodes_from(range(n))
        G.add_edges_from(zip(ei[0].tolist(), ei[1].tolist()))
        anchors = [int(a) for a in golden_small[f"micro/{name}/anchors"]]
        check(utils.all_pairs_shortest_path_length_parallel(G, anchors, workers), shims[f"{name}/all_pairs_keys/{workers}"],
              shims[f"{name}/all_pairs_rows/{workers}"], shims[f"{name}/all_pairs_is_int/{workers}"])
        part = shims[f"{name}/partition"]
        check(utils.shortest_path_length(G, anchors, part.tolist()), part, shims[f"{name}/partition_rows"],
              shims[f"{name}/partition_is_int"])
