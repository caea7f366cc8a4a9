"""CPU oracle for the GraphPOPE embedding-generation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  ``graphpope_b200`` never imports this package.

Parity status: the reference (JeroendenBoef/GraphPOPE) ships NO tests, golden
vectors or fixtures for this path, so the oracle is pinned
against outputs of the reference itself: ``tests/golden/generate_golden.py``
imports ``/root/reference/utils.py`` verbatim (through ``oracle/ref_shim.py``)
in the build container and freezes its outputs under ``tests/golden/``; the
CPU tests check every oracle function against those frozen vectors.

Modules
-------
geodesic   numpy/networkx restatement of utils.py:64-126 (+ to_networkx)
samplers   restatement of utils.py:18-62 (stochastic / degree / pagerank)
node2vec   restatement of utils.py:149-180 (pairwise + MinMaxScaler)
synth      synthetic graph generators for the BASELINE.json configs
cbfs       ctypes binding of bfs_oracle.c (plain C K x BFS, the fast tier)
ref_shim   container-only: import the real reference for fixture generation
"""
