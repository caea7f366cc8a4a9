"""Pin the CPU oracle (all tiers) to outputs frozen from the reference itself."""
import hashlib

import numpy as np
import pytest

from conftest import micro_names
from graphpope_b200 import synth
from oracle import cbfs, geodesic as g


def test_fp32_division_equals_reference_rounding():
    # utils.py:73,125: float64 1/len(path) rounded to float32 by torch.as_tensor.
    # The device epilogue computes IEEE fp32 1.0f/(float)(d+1); both must agree.
    n = np.arange(1, 65537, dtype=np.int64)
    via_f64 = (1.0 / n.astype(np.float64)).astype(np.float32)
    via_f32 = np.float32(1.0) / n.astype(np.float32)
    assert np.array_equal(via_f64.view(np.uint32), via_f32.view(np.uint32))


def test_known_answers_survey_8c(golden_small):
    rows = golden_small["micro/directed_chain/rows_f64"]
    assert np.allclose(rows[:, 0], [1 / 4, 1 / 3, 1 / 2, 1, 0])
    assert np.allclose(rows[:, 1], [1, 0, 0, 0, 0])


def test_all_tiers_match_reference_on_micro_graphs(golden_small):
    for name in micro_names(golden_small):
        n = int(golden_small[f"micro/{name}/n"])
        ei = golden_small[f"micro/{name}/edges"]
        anchors = golden_small[f"micro/{name}/anchors"]
        want = golden_small[f"micro/{name}/rows_f64"].astype(np.float32)
        t0 = g.t0_pairwise(ei, n, anchors)
        t1 = g.normalise_hops(g.t1_sssp_reverse_hops(ei, n, anchors))
        t2 = g.geodesic_features(ei, n, anchors)
        c = cbfs.geodesic_features(ei, n, anchors)
        for got in (t0, t1, t2, c):
            assert got.dtype == np.float32 and got.shape == want.shape, name
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name


def test_toy_pipeline_matches_reference(golden_small):
    n = int(golden_small["toy/n"])
    ei = golden_small["toy/edge_index"]
    anchors = golden_small["toy/anchors"]
    want = golden_small["toy/features"]
    assert np.array_equal(anchors, golden_small["samplers/stochastic_300_8"])
    assert list(anchors) == [102, 270, 106, 71, 188, 20, 102, 121]  # SURVEY §8c, duplicate 102
    x = golden_small["toy/x"]
    for feats in (g.geodesic_features(ei, n, anchors), cbfs.geodesic_features(ei, n, anchors)):
        got = g.concat_features(x, feats)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # direction matters on this asymmetric graph: forward BFS / symmetrised BFS differ
    assert not np.array_equal(g.geodesic_features(ei[::-1], n, anchors), want[:, x.shape[1]:])
    assert not np.array_equal(g.geodesic_features(ei, n, anchors, symmetrize=True), want[:, x.shape[1]:])


def test_toysym_two_lane_words(golden_small):
    ei = golden_small["toysym/edge_index"]
    anchors = golden_small["toysym/anchors"]
    want = golden_small["toysym/embedding"]
    assert anchors.size == 70
    got = cbfs.geodesic_features(ei, 300, anchors)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # symmetric input: symmetrisation is the identity
    assert np.array_equal(cbfs.geodesic_features(ei, 300, anchors, symmetrize=True), want)


def test_c_tier_matches_reference_on_full_pubmed_shape(golden_pubmed):
    """C1 of BASELINE.json: the reference verbatim (num_workers=6) vs the C oracle."""
    shape = synth.PUBMED_SHAPE
    ei = synth.make_graph(shape)
    anchors = synth.stochastic_anchors(shape.num_nodes, 256, seed=42)
    assert np.array_equal(anchors, golden_pubmed["anchors"])
    got = cbfs.geodesic_features(ei, shape.num_nodes, anchors)
    rows = golden_pubmed["sample_rows"]
    assert np.array_equal(got[rows].view(np.uint32), golden_pubmed["sample"].view(np.uint32))
    digest = hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest()
    assert digest == golden_pubmed["sha256"].tobytes()


def test_t2_equals_c_on_random_digraphs():
    for seed in range(6):
        n = 50 + 37 * seed
        ei = synth.random_digraph(n, 4 * n, seed=seed)
        anchors = np.random.default_rng(seed).integers(0, n, 9)
        a = g.t2_bfs_csr_hops(ei, n, anchors)
        b = cbfs.bfs_hops(cbfs.InCsr(ei, n), anchors)
        assert np.array_equal(a, b)
        assert np.array_equal(g.normalise_hops(a), cbfs.normalise(b))


def test_empty_and_degenerate_inputs():
    ei = np.zeros((2, 0), dtype=np.int64)
    assert g.geodesic_features(ei, 4, [1, 1]).tolist() == [[0, 0], [1, 1], [0, 0], [0, 0]]
    assert cbfs.geodesic_features(ei, 4, [1, 1]).tolist() == [[0, 0], [1, 1], [0, 0], [0, 0]]
    assert cbfs.geodesic_features(ei, 3, []).shape == (3, 0)
    with pytest.raises(ValueError):
        g.dedup_edges(np.array([[0], [5]]), 3)
    with pytest.raises(ValueError):
        cbfs.InCsr(np.array([[0], [5]]), 3)


def test_uint16_overflow_is_an_error_not_a_wrap():
    n = 65600  # path longer than the uint16 range
    ei = np.stack([np.arange(n - 1), np.arange(1, n)]).astype(np.int64)
    with pytest.raises(OverflowError):
        cbfs.bfs_hops(cbfs.InCsr(ei, n), [n - 1])
    ok = cbfs.bfs_hops(cbfs.InCsr(ei[:, : 65534], n), [65534])
    assert ok[0, 0] == 65534 and ok[65535, 0] == 0xFFFF


def test_node_slices_cover_range():
    # utils.py:100 float arithmetic; SURVEY App. A #6
    for n in (19717, 89250, 300):
        for w in (1, 2, 4, 6, 8):
            sl = g.node_slices(n, w)
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))


@pytest.mark.parametrize("name", ["deep_chain", "hub_star", "lollipop"])
def test_deep_and_hub_graphs_match_the_reference(golden_deep_hub, name):
    """4999-hop chain, 5000-leaf star + 300-leaf hub, lollipop with 70 anchors: what the unmodified reference's
    get_geodesic_distance_vector returned (tests/golden/generate_golden.py --deep-hub), bit for bit."""
    gd = golden_deep_hub
    n, ei, anchors = int(gd[f"{name}/n"]), gd[f"{name}/edge_index"], gd[f"{name}/anchors"]
    want = gd[f"{name}/embedding"]
    assert str(gd[f"{name}/dtype"]) == "torch.float32"
    hops = cbfs.bfs_hops(cbfs.InCsr(ei, n), anchors)
    got = cbfs.normalise(hops)
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(g.t2_bfs_csr_hops(ei, n, anchors), hops)
    deepest = {"deep_chain": 4999, "hub_star": 697, "lollipop": 45}[name]
    assert int(hops[hops != g.UNREACHABLE_U16].max()) == deepest
