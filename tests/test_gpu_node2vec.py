"""Parity of the tensor-core node2vec block (gp_cdist_minmax) with the reference / sklearn / torch.cdist.

north_star tolerance: 1e-4 relative vs torch.cdist in fp32; a pure relative test is meaningless at a
true zero (node == its own stochastic anchor), so an absolute floor of 1e-5 x (typical magnitude)
is added, as SURVEY.md §7 prescribes.
"""
import numpy as np
import pytest
import torch

from graphpope_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _close(got, want, scale):
    return np.allclose(got, want, rtol=RTOL, atol=1e-5 * scale)


@pytest.fixture(scope="module")
def dev():
    from graphpope_b200 import device
    return device


@pytest.mark.parametrize("fn", ["distance", "similarity", "euclidean"])
def test_block_matches_reference_fixture(dev, golden_small, fn):
    emb = synth.node2vec_table(400, 128, seed=int(golden_small["node2vec/table_seed"]))
    anchors = golden_small["node2vec/anchors"]
    want = golden_small[f"node2vec/{fn}"][:, 3:]  # reference attach_node2vec output minus the x columns
    got = dev.cdist_minmax(torch.as_tensor(emb), torch.as_tensor(emb[anchors]), fn, apply_minmax=True).cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.allclose(got, want, rtol=RTOL, atol=2e-5)  # values are in [0, 1] after scaling


@pytest.mark.parametrize("fn", ["distance", "similarity", "euclidean"])
@pytest.mark.parametrize("n,k,d", [(1000, 12, 128), (257, 256, 128), (5000, 300, 128), (640, 64, 64), (1, 1, 128)])
def test_raw_pairwise_matches_oracle_and_torch(dev, fn, n, k, d):
    from oracle import node2vec as nv
    rng = np.random.default_rng(n + k)
    emb = rng.standard_normal((n, d)).astype(np.float32)
    idx = rng.integers(0, n, k)
    idx[: min(k, 3)] = idx[0]  # duplicate anchors -> duplicate columns, several exact self distances
    anc = emb[idx].copy()
    got = dev.cdist_minmax(torch.as_tensor(emb), torch.as_tensor(anc), fn, apply_minmax=False).cpu().numpy()
    want = nv.PAIRWISE[fn](emb, anc)
    scale = float(np.sqrt(2 * d)) if fn == "euclidean" else 1.0
    assert _close(got, want, scale), np.abs(got - want).max()
    if fn == "euclidean":
        ref = torch.cdist(torch.as_tensor(emb), torch.as_tensor(anc),
                          compute_mode="donot_use_mm_for_euclid_dist").numpy()
        assert _close(got, ref, scale)
        assert np.all(got[idx, np.arange(k)] < 1e-5 * scale)  # node == its own anchor: (near) zero, as torch gives
    assert np.array_equal(got[:, 0], got[:, min(k, 3) - 1])


@pytest.mark.parametrize("fn", ["distance", "euclidean"])
def test_minmax_columns(dev, fn):
    from oracle import node2vec as nv
    emb = synth.node2vec_table(3000, 128, seed=5)
    anc = emb[[5, 5, 77, 2999]].copy()
    out = torch.full((3000, 10), -7.0, device="cuda")
    dev.cdist_minmax(torch.as_tensor(emb), torch.as_tensor(anc), fn, True, out=out, col_offset=4)
    got = out.cpu().numpy()
    want = nv.node2vec_block(emb, anc, fn)
    assert np.all(got[:, :4] == -7.0) and np.all(got[:, 8:] == -7.0)
    assert np.allclose(got[:, 4:8], want, rtol=RTOL, atol=2e-5)
    assert got[:, 4:8].min() >= -1e-6 and got[:, 4:8].max() <= 1 + 1e-6


def test_flickr_shape_config_c3(dev):
    """BASELINE.json configs[2]: 89,250 x 128 table, 256 anchors; linearity/permutation properties."""
    from oracle import node2vec as nv
    emb = synth.node2vec_table(89250, 128, seed=3)
    idx = synth.stochastic_anchors(89250, 256, 42)
    e_d, a_d = torch.as_tensor(emb).cuda(), torch.as_tensor(emb[idx]).cuda()
    got = dev.cdist_minmax(e_d, a_d, "euclidean", apply_minmax=False)
    rows = np.arange(0, 89250, 41)
    assert _close(got[rows].cpu().numpy(), nv.euclidean_distances(emb[rows], emb[idx]), 16.0)
    perm = torch.randperm(256)
    got_p = dev.cdist_minmax(e_d, a_d[perm.cuda()], "euclidean", apply_minmax=False)
    assert torch.equal(got_p, got[:, perm.cuda()])  # anchor order only permutes columns (bit-exact)
    scaled = dev.cdist_minmax(e_d, a_d, "similarity", apply_minmax=True)
    assert float(scaled.min()) == 0.0 and abs(float(scaled.max()) - 1.0) < 1e-6


def test_attach_node2vec_end_to_end(dev, golden_small, tmp_path, monkeypatch):
    from graphpope_b200 import utils
    monkeypatch.setenv("GRAPHPOPE_DATA_DIR", str(tmp_path))
    table = synth.node2vec_table(400, 128, seed=int(golden_small["node2vec/table_seed"]))
    torch.save(torch.as_tensor(table), tmp_path / "toy_node2vec.pt")

    class Data:
        pass

    for fn in ("distance", "similarity", "euclidean"):
        d = Data()
        d.num_nodes, d.edge_index, d.x = 400, None, torch.as_tensor(golden_small["node2vec/x"])
        np.random.seed(5)
        got = utils.attach_node2vec(d, "toy", 12, "stochastic", fn, 2)
        want = golden_small[f"node2vec/{fn}"]
        assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
        assert np.allclose(got.numpy(), want, rtol=RTOL, atol=2e-5)


def test_bad_arguments_are_loud(dev):
    from graphpope_b200._lib import GraphpopeError
    with pytest.raises(GraphpopeError):
        dev.cdist_minmax(torch.zeros(10, 100), torch.zeros(2, 100), 7)  # unknown mode
    with pytest.raises(ValueError):
        dev.cdist_minmax(torch.zeros(10, 100), torch.zeros(2, 64), "euclidean")  # widths differ
    with pytest.raises(ValueError):
        dev.kmeans(torch.zeros(300, 200), 4)  # device k-means covers widths up to 128


def test_kmeans_assign_matches_fp32_argmin():
    """The tcgen05 kernel in arg-min mode: nearest centre and squared distance of every row."""
    import ctypes
    from graphpope_b200 import _lib
    from graphpope_b200.device import _ptr, _stream
    lib = _lib.require_cuda()
    rng = np.random.default_rng(5)
    for n, k, d in ((1000, 7, 64), (5000, 256, 128), (3000, 300, 128)):
        x = torch.as_tensor(rng.standard_normal((n, d)).astype(np.float32)).cuda()
        c = torch.as_tensor(rng.standard_normal((k, d)).astype(np.float32)).cuda()
        best = torch.empty(n, dtype=torch.int64, device="cuda")
        _lib.check(lib.gp_kmeans_assign(_ptr(x), _ptr(c), n, k, d, _ptr(best), _stream()))
        lab = (best & 0xFFFFFFFF).cpu().numpy()
        d2 = (best >> 32).to(torch.int32).view(torch.float32).cpu().numpy()
        full = torch.cdist(x.double(), c.double()).pow(2).cpu().numpy()
        want = full.min(axis=1)
        assert np.allclose(d2, want, rtol=1e-4, atol=1e-4)
        # the chosen centre is (numerically) a nearest one
        assert np.all(full[np.arange(n), lab] <= want * (1 + 1e-4) + 1e-4)


def test_device_kmeans_inertia_close_to_sklearn():
    """Statistical parity (the reference runs scikit-learn unseeded, utils.py:169): same objective within
    2 % on a clustered table and on the Gaussian table the reference's generator actually produces."""
    from sklearn.cluster import KMeans
    from graphpope_b200 import device as dev
    rng = np.random.default_rng(11)
    centres = rng.standard_normal((24, 64)).astype(np.float32) * 4
    clustered = (centres[rng.integers(0, 24, 6000)] + rng.standard_normal((6000, 64)).astype(np.float32)).astype(np.float32)
    gauss = synth.node2vec_table(4000, 128, seed=3)
    for table, k in ((clustered, 24), (gauss, 32)):
        got_c, got_inertia, iters = dev.kmeans(torch.as_tensor(table), k, n_init=4, seed=0)
        ref = KMeans(n_clusters=k, n_init=4, random_state=0).fit(table)
        assert got_c.shape == (k, table.shape[1]) and 1 <= iters <= 300
        assert got_inertia <= ref.inertia_ * 1.02, (got_inertia, ref.inertia_)
        # the reported inertia is the objective of the returned centres
        d2 = torch.cdist(torch.as_tensor(table).double(), got_c.cpu().double()).pow(2).min(dim=1).values.sum().item()
        assert abs(d2 - got_inertia) <= 1e-3 * d2


@pytest.mark.parametrize("fn", ["distance", "similarity", "euclidean"])
@pytest.mark.parametrize("n,k,d", [(700, 40, 100), (513, 70, 32), (300, 33, 7), (900, 80, 200), (257, 300, 130)])
def test_any_embedding_width(dev, fn, n, k, d):
    """utils.py:174 accepts any embedding width: widths below 128 are zero-padded onto the tensor-core kernel,
    wider tables take the plain fp32 kernel; raw and MinMax-scaled blocks against the sklearn restatement."""
    from oracle import node2vec as nv
    rng = np.random.default_rng(n + k + d)
    emb = rng.standard_normal((n, d)).astype(np.float32)
    idx = rng.integers(0, n, k)
    anc = emb[idx].copy()
    scale = float(np.sqrt(2 * d)) if fn == "euclidean" else 1.0
    got = dev.cdist_minmax(torch.as_tensor(emb), torch.as_tensor(anc), fn, apply_minmax=False).cpu().numpy()
    want = nv.PAIRWISE[fn](emb, anc)
    # the dropped lo*lo term of the split-bf16 product is ~2^-16 of |x||a| whatever the width; short rows do not
    # average it down, so the absolute floor is wider for them (the relative bar of north_star stays 1e-4)
    floor = (1e-5 if d >= 32 else 5e-5) * scale
    assert np.allclose(got, want, rtol=RTOL, atol=floor), np.abs(got - want).max()
    got = dev.cdist_minmax(torch.as_tensor(emb), torch.as_tensor(anc), fn, apply_minmax=True).cpu().numpy()
    lo, hi = want.min(axis=0), want.max(axis=0)
    rng_ = np.where(hi - lo < 10 * np.finfo(np.float32).eps, 1.0, hi - lo)
    assert np.allclose(got, (want - lo) / rng_, rtol=RTOL, atol=3e-5 if d >= 32 else 1e-4), np.abs(got - (want - lo) / rng_).max()


def test_kmeans_on_a_100_column_table(dev):
    """Device k-means pads a 100-column table to 128; centres come back 100 wide and cluster a separable table."""
    rng = np.random.default_rng(0)
    centres = rng.standard_normal((8, 100)).astype(np.float32) * 6
    x = (centres[rng.integers(0, 8, 4000)] + rng.standard_normal((4000, 100)).astype(np.float32) * 0.3)
    got, inertia, _ = dev.kmeans(torch.as_tensor(x), 8, n_init=3, seed=1)
    assert tuple(got.shape) == (8, 100)
    d = torch.cdist(got.cpu(), torch.as_tensor(centres))
    assert float(d.min(dim=1).values.max()) < 0.2 and inertia < 4000 * 100 * 0.3 ** 2 * 1.2
