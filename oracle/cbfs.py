"""ctypes binding of bfs_oracle.c (plain-C K x BFS).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbfs_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bfs_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.gpo_build_in_csr.restype = ctypes.c_int
        L.gpo_build_in_csr.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), i64p]
        L.gpo_free.restype = None
        L.gpo_free.argtypes = [ctypes.c_void_p]
        L.gpo_bfs_hops.restype = ctypes.c_int64
        L.gpo_bfs_hops.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                   ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
        L.gpo_normalise.restype = None
        L.gpo_normalise.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _lib = L
    return _lib


class InCsr:
    """In-edge CSR held in C memory; ``rowptr`` / ``col`` are numpy views."""

    def __init__(self, edge_index, num_nodes: int, symmetrize: bool = False):
        ei = np.ascontiguousarray(np.asarray(edge_index, dtype=np.int64).reshape(2, -1))
        rp, cp, eu = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        rc = lib().gpo_build_in_csr(ei.ctypes.data, ei.shape[1], int(num_nodes), int(symmetrize),
                                    ctypes.byref(rp), ctypes.byref(cp), ctypes.byref(eu))
        if rc != 0:
            raise ValueError("gpo_build_in_csr failed (index outside [0, N) or out of memory)")
        self.num_nodes = int(num_nodes)
        self.num_edges = int(eu.value)
        self._rp, self._cp = rp, cp
        self.rowptr = np.ctypeslib.as_array(ctypes.cast(rp, ctypes.POINTER(ctypes.c_int64)),
                                            shape=(self.num_nodes + 1,))
        self.col = np.ctypeslib.as_array(ctypes.cast(cp, ctypes.POINTER(ctypes.c_int32)),
                                         shape=(max(self.num_edges, 1),))[: self.num_edges]

    def __del__(self):
        try:
            lib().gpo_free(self._rp)
            lib().gpo_free(self._cp)
        except Exception:
            pass


def bfs_hops(csr: InCsr, anchors, out: np.ndarray | None = None, col_offset: int = 0) -> np.ndarray:
    """uint16 ``[N, K]`` hop matrix (0xFFFF = unreachable)."""
    a = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64))
    if out is None:
        out = np.empty((csr.num_nodes, a.size), dtype=np.uint16)
    ld = out.shape[1] if out.ndim == 2 else a.size
    rc = lib().gpo_bfs_hops(csr._rp, csr._cp, csr.num_nodes, a.ctypes.data, a.size,
                            out.ctypes.data, ld, col_offset)
    if rc == -2:
        raise OverflowError("hop distance does not fit uint16")
    if rc < 0:
        raise ValueError("gpo_bfs_hops failed")
    bfs_hops.last_max_level = int(rc)
    return out


def normalise(dist_u16: np.ndarray) -> np.ndarray:
    d = np.ascontiguousarray(dist_u16, dtype=np.uint16)
    out = np.empty(d.shape, dtype=np.float32)
    lib().gpo_normalise(d.ctypes.data, d.size, out.ctypes.data)
    return out


def geodesic_features(edge_index, num_nodes, anchors, symmetrize=False) -> np.ndarray:
    return normalise(bfs_hops(InCsr(edge_index, num_nodes, symmetrize), anchors))
