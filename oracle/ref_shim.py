"""Import the UNMODIFIED reference ``utils.py`` from /root/reference.

Container-only helper (``/root/reference`` does not exist on the GPU box): used
by ``tests/golden/generate_golden.py`` to freeze reference outputs as fixtures
and by the optional ``-m "not gpu"`` cross-checks that skip when the reference
is absent.  Test infrastructure, never imported by the product.

The reference fails to import only on ``from torch_geometric.utils import
to_networkx`` (utils.py:12; PyG is not installed and there is no network), so a
stub ``torch_geometric.utils`` module is injected that restates PyG 1.7's
documented ``to_networkx`` defaults (``to_undirected=False``): a ``DiGraph``
with nodes ``0..num_nodes-1`` and one ``add_edge(u, v)`` per ``edge_index``
column (parallel edges collapse, self-loops stay).  ``nx.pagerank_scipy`` was
removed in networkx 3; it is aliased to ``nx.pagerank`` (the same scipy power
iteration) so utils.py:28 runs.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_DIR = "/root/reference"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "utils.py"))


def _to_networkx(data, node_attrs=None, edge_attrs=None, to_undirected=False,
                 remove_self_loops=False):
    import networkx as nx

    G = nx.Graph() if to_undirected else nx.DiGraph()
    G.add_nodes_from(range(int(data.num_nodes)))
    ei = data.edge_index
    rows = ei[0].tolist()
    cols = ei[1].tolist()
    for u, v in zip(rows, cols):
        if to_undirected and v > u:
            continue
        if remove_self_loops and u == v:
            continue
        G.add_edge(u, v)
    return G


def load_reference_utils():
    """Return the reference's ``utils`` module object (imported verbatim)."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    import networkx as nx

    if "torch_geometric" not in sys.modules:
        pkg = types.ModuleType("torch_geometric")
        sub = types.ModuleType("torch_geometric.utils")
        sub.to_networkx = _to_networkx
        pkg.utils = sub
        sys.modules["torch_geometric"] = pkg
        sys.modules["torch_geometric.utils"] = sub
    if not hasattr(nx, "pagerank_scipy"):
        nx.pagerank_scipy = nx.pagerank
    name = "_graphpope_reference_utils"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(REFERENCE_DIR, "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class RefData:
    """Duck-typed stand-in for a PyG ``Data`` object."""

    def __init__(self, edge_index, num_nodes, x=None):
        self.edge_index = edge_index
        self.num_nodes = int(num_nodes)
        self.x = x
