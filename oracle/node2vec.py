"""CPU restatement of the GraphPOPE-node2vec block (reference utils.py:158-176).

TEST INFRASTRUCTURE — see oracle/__init__.py.

Third-party arithmetic restated: scikit-learn (pinned 0.24.2 by the reference's
requirements.txt:4; 1.9.0 installed — the pairwise and scaler formulas are the
same) ``cosine_similarity``, ``cosine_distances``, ``euclidean_distances`` and
``MinMaxScaler``; call sites utils.py:158-162,174-176.  tests pin these
restatements against the installed sklearn functions themselves.
"""
from __future__ import annotations

import numpy as np

MODES = ("distance", "similarity", "euclidean")  # keys of dist_map, utils.py:158-162


def _row_normalise(a: np.ndarray) -> np.ndarray:
    norms = np.sqrt(np.einsum("ij,ij->i", a, a))
    norms = np.where(norms == 0.0, 1.0, norms).astype(a.dtype)
    return a / norms[:, None]


def cosine_similarity(x: np.ndarray, a: np.ndarray) -> np.ndarray:
    """sklearn: normalise rows (input dtype kept), then one GEMM."""
    return _row_normalise(x) @ _row_normalise(a).T


def cosine_distances(x: np.ndarray, a: np.ndarray) -> np.ndarray:
    """sklearn: ``1 - cosine_similarity`` clipped to [0, 2]."""
    s = cosine_similarity(x, a)
    s *= -1
    s += 1
    np.clip(s, 0, 2, out=s)
    return s


def euclidean_distances(x: np.ndarray, a: np.ndarray) -> np.ndarray:
    """sklearn float32 path: the expansion is evaluated in float64
    (``_euclidean_distances_upcast``), cast back, clamped at 0, then sqrt."""
    x64 = x.astype(np.float64)
    a64 = a.astype(np.float64)
    d = -2.0 * (x64 @ a64.T)
    d += np.einsum("ij,ij->i", x64, x64)[:, None]
    d += np.einsum("ij,ij->i", a64, a64)[None, :]
    d = d.astype(x.dtype if x.dtype == np.float32 else np.float64)
    np.maximum(d, 0, out=d)
    return np.sqrt(d, out=d)


PAIRWISE = {
    "distance": cosine_distances,
    "similarity": cosine_similarity,
    "euclidean": euclidean_distances,
}


def minmax_scale_columns(m: np.ndarray) -> np.ndarray:
    """``MinMaxScaler().fit(m).transform(m)``: per COLUMN over all rows.

    scale = 1/range with range < 10*eps replaced by 1 (``_handle_zeros_in_scale``),
    min_ = -data_min*scale, out = m*scale + min_; all in the input dtype.
    """
    m = np.asarray(m)
    if m.shape[0] == 0:
        raise ValueError("MinMaxScaler needs at least one sample")
    dmin = m.min(axis=0)
    rng = m.max(axis=0) - dmin
    rng = np.where(rng < 10 * np.finfo(m.dtype).eps, 1.0, rng).astype(m.dtype)
    scale = (1.0 / rng).astype(m.dtype)
    mn = (0 - dmin * scale).astype(m.dtype)
    return (m * scale + mn).astype(m.dtype)


def node2vec_block(emb: np.ndarray, anchor_emb: np.ndarray, distance_function: str) -> np.ndarray:
    """utils.py:174-176: pairwise then per-column min-max; float32 ``[N, K]``."""
    fn = PAIRWISE[distance_function]  # KeyError on an unknown key, as utils.py:164
    return minmax_scale_columns(fn(np.asarray(emb, dtype=np.float32),
                                   np.asarray(anchor_emb, dtype=np.float32)))
