"""The reference-shaped pool helpers (utils.py:64-114) on the GPU path: key order, Python value types and
partition subsets against what the UNMODIFIED reference returned (tests/golden/reference_shims.npz, frozen by
tests/golden/generate_golden.py --shims) and against the micro rows of reference_small.npz."""
import os

import networkx as nx
import numpy as np
import pytest

from conftest import GOLDEN_DIR, micro_names

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden_shims():
    return np.load(os.path.join(GOLDEN_DIR, "reference_shims.npz"))


def _graph(golden_small, name):
    n = int(golden_small[f"micro/{name}/n"])
    ei = golden_small[f"micro/{name}/edges"]
    G = nx.DiGraph()
    G.add_nodes_from(range(n))
    G.add_edges_from(zip(ei[0].tolist(), ei[1].tolist()))
    return G, n, [int(a) for a in golden_small[f"micro/{name}/anchors"]]


def _check(got, keys, rows, is_int):
    assert list(got.keys()) == keys.tolist()
    for k, want_row, int_row in zip(keys.tolist(), rows, is_int):
        row = got[k]
        assert isinstance(row, list) and len(row) == len(want_row)
        for v, w, isint in zip(row, want_row, int_row):
            assert type(v) is (int if isint else float), (k, v)
            assert v == w  # float64 1/len(path) is exact in both


def test_shortest_path_length_values_types_and_order(golden_small, golden_shims):
    from graphpope_b200 import utils
    for name in micro_names(golden_small):
        G, n, anchors = _graph(golden_small, name)
        got = utils.shortest_path_length(G, anchors, list(range(n)))
        want = golden_small[f"micro/{name}/rows_f64"]
        assert list(got.keys()) == list(range(n))
        for i in range(n):
            assert got[i] == want[i].tolist()
            assert all(type(v) is (int if w == 0 else float) for v, w in zip(got[i], want[i]))
        part = golden_shims[f"{name}/partition"]
        got = utils.shortest_path_length(G, anchors, part.tolist())
        _check(got, part, golden_shims[f"{name}/partition_rows"], golden_shims[f"{name}/partition_is_int"])


@pytest.mark.parametrize("workers", [1, 2, 3])
def test_all_pairs_parallel_matches_reference_dict(golden_small, golden_shims, workers):
    from graphpope_b200 import utils
    for name in micro_names(golden_small):
        G, n, anchors = _graph(golden_small, name)
        got = utils.all_pairs_shortest_path_length_parallel(G, anchors, workers)
        _check(got, golden_shims[f"{name}/all_pairs_keys/{workers}"], golden_shims[f"{name}/all_pairs_rows/{workers}"],
               golden_shims[f"{name}/all_pairs_is_int/{workers}"])


def test_get_geodesic_distance_vector_matches_dict_path(golden_small):
    """torch.as_tensor(list(dist_dict.values())) (utils.py:125) of the dict path == the fused device path."""
    import torch
    from graphpope_b200 import utils
    name = "two_components"
    G, n, anchors = _graph(golden_small, name)
    d = utils.all_pairs_shortest_path_length_parallel(G, anchors, 2)
    via_dict = torch.as_tensor(list(d.values()))

    class Data:
        pass

    data = Data()
    data.num_nodes, data.edge_index = n, torch.as_tensor(golden_small[f"micro/{name}/edges"])
    data.anchor_nodes = anchors
    direct = utils.get_geodesic_distance_vector(data, 2)
    assert direct.dtype == torch.float32 and torch.equal(direct, via_dict.to(torch.float32))


def test_lollipop_deep_hub_two_lane_words_matches_reference():
    """reference_deep_hub.npz/lollipop: 45 hops (deep bit planes), a 201-edge row (hub chunks) and 70 anchors with
    repeats (two lane words, ragged decode) in one graph, against the unmodified reference's
    get_geodesic_distance_vector, bit for bit, through the host entry and through the device engine."""
    import torch
    from graphpope_b200 import device as dev
    g = np.load(os.path.join(GOLDEN_DIR, "reference_deep_hub.npz"))
    n, ei, anchors, want = int(g["lollipop/n"]), g["lollipop/edge_index"], g["lollipop/anchors"], g["lollipop/embedding"]
    out, _, stats = dev.geodesic_embed_host(torch.as_tensor(ei), n, anchors, None, False)
    assert stats["max_level"] == 45
    assert np.array_equal(out.numpy().view(np.uint32), want.view(np.uint32))
    eng = dev.GeodesicEngine(n, ei.shape[1], len(anchors))
    feats = eng.run(torch.as_tensor(ei).cuda(), torch.as_tensor(anchors).cuda(), None)
    assert np.array_equal(feats.cpu().numpy().view(np.uint32), want.view(np.uint32))
