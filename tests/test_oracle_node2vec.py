"""Pin the node2vec-branch restatement to the reference and to sklearn."""
import numpy as np
import pytest

from graphpope_b200 import synth
from oracle import node2vec as nv


def _table(golden_small, n=400):
    return synth.node2vec_table(n, 128, seed=int(golden_small["node2vec/table_seed"]))


@pytest.mark.parametrize("fn", nv.MODES)
def test_block_matches_reference(golden_small, fn):
    emb = _table(golden_small)
    anchors = golden_small["node2vec/anchors"]
    x = golden_small["node2vec/x"]
    want = golden_small[f"node2vec/{fn}"]
    got = np.concatenate([x, nv.node2vec_block(emb, emb[anchors], fn)], axis=1)
    assert got.dtype == np.float32 and got.shape == want.shape
    # BLAS summation order is not pinned; values live in [0, 1] after scaling
    assert np.allclose(got, want, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("fn", nv.MODES)
def test_pairwise_matches_sklearn(fn):
    from sklearn.metrics import pairwise as skp
    from sklearn.preprocessing import MinMaxScaler

    sk = {"distance": skp.cosine_distances, "similarity": skp.cosine_similarity,
          "euclidean": skp.euclidean_distances}[fn]
    emb = synth.node2vec_table(700, 128, seed=9)
    a = emb[[3, 3, 10, 699, 0]]
    raw = sk(emb, a)
    assert np.allclose(nv.PAIRWISE[fn](emb, a), raw, rtol=1e-5, atol=2e-6)
    want = MinMaxScaler().fit(raw).transform(raw)
    assert np.allclose(nv.minmax_scale_columns(raw), want, rtol=0, atol=1e-7)


def test_constant_column_scale_is_one():
    m = np.array([[2.0, 1.0], [2.0, 3.0]], dtype=np.float32)
    assert nv.minmax_scale_columns(m).tolist() == [[0.0, 0.0], [0.0, 1.0]]


def test_unknown_distance_function_raises_keyerror():
    with pytest.raises(KeyError):
        nv.node2vec_block(np.zeros((2, 4), np.float32), np.zeros((1, 4), np.float32), "None")
