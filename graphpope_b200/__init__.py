"""graphpope_b200 — B200-native (sm_100a) GraphPOPE embedding generation.

The reference-facing API lives in :mod:`graphpope_b200.utils` and mirrors the
reference's ``utils.py`` name for name (``Graphpope``, ``attach_distance_embedding``,
``attach_node2vec``, ``sample_anchor_nodes``, ``get_geodesic_distance_vector``,
``concat_into_features``, ...).  :mod:`graphpope_b200.device` is the tensor-level
layer over the C ABI (``include/graphpope_b200.h``).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
