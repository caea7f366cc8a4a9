"""Freeze outputs of the UNMODIFIED reference as golden fixtures.

Run in the build container only (needs /root/reference):

    python tests/golden/generate_golden.py            # small fixtures, seconds
    python tests/golden/generate_golden.py --pubmed   # + full C1 config (~2 min, 6 workers)
    python tests/golden/generate_golden.py --betweenness   # ONLY reference_betweenness.npz (seconds)
    python tests/golden/generate_golden.py --eigenvector   # ONLY reference_eigenvector.npz (seconds)
    python tests/golden/generate_golden.py --shims         # ONLY reference_shims.npz (seconds)
    python tests/golden/generate_golden.py --deep-hub      # ONLY reference_deep_hub.npz (~1 min)

The reference has no tests or golden vectors of its own, so these
files ARE the parity pins: every value below is produced by
/root/reference/utils.py itself, imported verbatim through oracle/ref_shim.py.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from graphpope_b200 import synth  # noqa: E402
from oracle.ref_shim import RefData, load_reference_utils  # noqa: E402


def micro_graphs():
    """name -> (num_nodes, edge list, anchors)."""
    return {
        # SURVEY §8c known answers: 0->1->2->3, isolated 4
        "directed_chain": (5, [(0, 1), (1, 2), (2, 3)], [3, 0]),
        "path_sym": (6, [(i, i + 1) for i in range(5)] + [(i + 1, i) for i in range(5)], [0, 5, 2]),
        "star_sym": (7, [(0, i) for i in range(1, 7)] + [(i, 0) for i in range(1, 7)], [0, 3]),
        "two_components": (8, [(0, 1), (1, 0), (1, 2), (2, 1), (4, 5), (5, 4), (5, 6), (6, 5), (6, 7), (7, 6)],
                           [0, 7, 3]),
        "self_loop": (4, [(0, 0), (0, 1), (1, 2), (2, 2), (3, 3)], [2, 3, 0]),
        "duplicate_edge": (4, [(0, 1), (0, 1), (1, 2), (1, 2), (2, 3), (0, 1)], [3, 1]),
        "isolated_and_dup_anchor": (5, [(0, 1), (1, 0), (2, 3)], [1, 1, 4, 3, 1]),
        "in_star_directed": (5, [(1, 0), (2, 0), (3, 0), (4, 0)], [0, 1]),
        "out_star_directed": (5, [(0, 1), (0, 2), (0, 3), (0, 4)], [0, 1]),
        "cycle_directed": (5, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0)], [0, 2]),
    }


def ref_rows(utils, n, edges, anchors):
    ei = torch.tensor(np.asarray(edges, dtype=np.int64).reshape(-1, 2).T.copy())
    data = RefData(ei, n)
    G = utils.to_networkx(data)
    rows = utils.shortest_path_length(G, anchors, list(range(n)))
    return np.asarray([rows[i] for i in range(n)], dtype=np.float64)


def betweenness_fixture(utils):
    """reference_betweenness.npz: utils.sample_anchor_nodes(..., 'betweenness_centrality') (utils.py:32-36) and
    the nx.betweenness_centrality scores it ranks, on a 600-node graph with a symmetric heavy-tailed part and
    an asymmetric part (so in- and out-neighbourhoods differ)."""
    import networkx as nx

    n = 600
    ei = synth.chung_lu_symmetric(n, 3600, 2.2, seed=41)
    ei = np.concatenate([ei, synth.random_digraph(n, 150, seed=42)], axis=1)
    data = RefData(torch.tensor(ei), n)
    out = {"n": np.int64(n), "edge_index": ei}
    for k in (1, 16, 64, 256):
        out[f"anchors/{k}"] = np.asarray(utils.sample_anchor_nodes(data, k, "betweenness_centrality"), dtype=np.int64)
    bc = nx.betweenness_centrality(utils.to_networkx(data))
    out["scores"] = np.asarray([bc[i] for i in range(n)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "reference_betweenness.npz"), **out)
    print("wrote reference_betweenness.npz; non-zero scores:", int((out["scores"] > 0).sum()))


def strongly_connected_graph(n_target=800, seed=51):
    """Giant component of a heavy-tailed symmetric graph, relabelled 0..m-1, plus random one-way edges
    (adding edges keeps it strongly connected; they make in- and out-neighbourhoods differ)."""
    import networkx as nx

    ei = synth.chung_lu_symmetric(n_target, 6 * n_target, 2.2, seed=seed)
    G = nx.Graph()
    G.add_nodes_from(range(n_target))
    G.add_edges_from(zip(ei[0].tolist(), ei[1].tolist()))
    giant = sorted(max(nx.connected_components(G), key=len))
    relabel = {v: i for i, v in enumerate(giant)}
    keep = [(relabel[u], relabel[v]) for u, v in zip(ei[0].tolist(), ei[1].tolist()) if u in relabel and v in relabel]
    m = len(giant)
    extra = synth.random_digraph(m, m // 4, seed=seed + 1)
    out = np.concatenate([np.asarray(keep, dtype=np.int64).T, extra], axis=1)
    return np.ascontiguousarray(out), m


def eigenvector_fixture(utils):
    """reference_eigenvector.npz: utils.sample_anchor_nodes(..., 'eigenvector_centrality') (utils.py:44-48) and
    the nx.eigenvector_centrality_numpy scores it ranks, on a strongly connected digraph (networkx >= 3.2
    refuses anything else)."""
    import networkx as nx

    ei, n = strongly_connected_graph()
    data = RefData(torch.tensor(ei), n)
    out = {"n": np.int64(n), "edge_index": ei}
    for k in (1, 16, 64, 256):
        out[f"anchors/{k}"] = np.asarray(utils.sample_anchor_nodes(data, k, "eigenvector_centrality"), dtype=np.int64)
    ev = nx.eigenvector_centrality_numpy(utils.to_networkx(data))
    out["scores"] = np.asarray([ev[i] for i in range(n)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "reference_eigenvector.npz"), **out)
    print("wrote reference_eigenvector.npz; n =", n, "min score", float(out["scores"].min()))


def shims_fixture(utils):
    """reference_shims.npz: what utils.shortest_path_length (utils.py:64-81) and
    utils.all_pairs_shortest_path_length_parallel (utils.py:92-114) return as Python objects — key order,
    the Python type of every entry (float 1/len(path) vs the int 0 appended at utils.py:76) and the rows of a
    partition that is a permuted subset of the nodes."""
    out = {}
    for name, (n, edges, anchors) in micro_graphs().items():
        ei = torch.tensor(np.asarray(edges, dtype=np.int64).reshape(-1, 2).T.copy())
        G = utils.to_networkx(RefData(ei, n))
        part = list(range(n))[::-2]  # descending, every other node
        rows = utils.shortest_path_length(G, anchors, part)
        out[f"{name}/partition"] = np.asarray(list(rows.keys()), dtype=np.int64)
        out[f"{name}/partition_rows"] = np.asarray([rows[k] for k in rows], dtype=np.float64)
        out[f"{name}/partition_is_int"] = np.asarray([[isinstance(v, int) for v in rows[k]] for k in rows], dtype=bool)
        for w in (1, 2, 3):
            full = utils.all_pairs_shortest_path_length_parallel(G, anchors, w)
            out[f"{name}/all_pairs_keys/{w}"] = np.asarray(list(full.keys()), dtype=np.int64)
            out[f"{name}/all_pairs_rows/{w}"] = np.asarray([full[k] for k in full], dtype=np.float64)
            out[f"{name}/all_pairs_is_int/{w}"] = np.asarray([[isinstance(v, int) for v in full[k]] for k in full],
                                                             dtype=bool)
    np.savez_compressed(os.path.join(HERE, "reference_shims.npz"), **out)
    print("wrote reference_shims.npz with", len(out), "arrays")


def deep_hub_graphs():
    """name -> (num_nodes, edge_index [2,E], anchors): the shapes that take the special paths of the device kernels —
    hop counts far beyond the 15 one-hot result arrays (deep bit planes), rows beyond 128 edges (hub chunks, CTA row
    sort), and a mix of both with more than 64 anchors (two lane words, repeated anchors)."""
    out = {}
    n = 5000  # the graph of tests/test_gpu_geodesic.py::test_deep_path_uses_many_distance_planes
    out["deep_chain"] = (n, np.stack([np.arange(n - 1), np.arange(1, n)]).astype(np.int64),
                         np.array([n - 1, 0, 2500], dtype=np.int64))
    n = 6000  # the graph of test_hub_rows_take_the_cta_and_warp_paths
    e = [(0, i) for i in range(1, 5001)] + [(i, 0) for i in range(1, 5001)]
    e += [(5001, i) for i in range(5002, 5302)] + [(i, 5001) for i in range(5002, 5302)]
    e += [(i, i + 1) for i in range(5302, 5999)] + [(5001, 0), (5500, 5001)]
    out["hub_star"] = (n, np.asarray(e, dtype=np.int64).T, np.array([0, 17, 5001, 5999, 5302, 5100, 17], dtype=np.int64))
    # lollipop: a two-way chain of 60 nodes whose last node is a two-way hub of 200 leaves, one-way shortcuts, an
    # isolated node; 70 anchors with repeats
    n = 60 + 200 + 1
    e = [(i, i + 1) for i in range(59)] + [(i + 1, i) for i in range(59)]
    e += [(59, j) for j in range(60, 260)] + [(j, 59) for j in range(60, 260)]
    e += [(0, 30), (100, 5), (200, 201), (201, 200), (7, 7), (0, 1)]
    rng = np.random.default_rng(77)
    anchors = np.concatenate([[0, 59, 60, 259, 260, 30, 30], rng.integers(0, n, 63)]).astype(np.int64)
    out["lollipop"] = (n, np.asarray(e, dtype=np.int64).T, anchors)
    return out


def deep_hub_fixture(utils):
    """reference_deep_hub.npz: utils.get_geodesic_distance_vector (utils.py:116-126, pool of 2 included) of the
    unmodified reference on the graphs above (about a minute: N*K nx.shortest_path calls on a 5000-hop chain)."""
    out = {}
    for name, (n, ei, anchors) in deep_hub_graphs().items():
        data = RefData(torch.tensor(ei), n)
        data.anchor_nodes = anchors
        emb = utils.get_geodesic_distance_vector(data, 2)
        assert tuple(emb.shape) == (n, anchors.size)
        out[f"{name}/n"] = np.int64(n)
        out[f"{name}/edge_index"] = ei
        out[f"{name}/anchors"] = anchors
        out[f"{name}/embedding"] = emb.numpy().astype(np.float32)
        out[f"{name}/dtype"] = np.asarray(str(emb.dtype))
        print(name, "max hops", int(round(1.0 / float(emb[emb > 0].min()) - 1)), "dtype", emb.dtype)
    np.savez_compressed(os.path.join(HERE, "reference_deep_hub.npz"), **out)
    print("wrote reference_deep_hub.npz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--deep-hub", action="store_true")
    ap.add_argument("--pubmed", action="store_true")
    ap.add_argument("--betweenness", action="store_true")
    ap.add_argument("--eigenvector", action="store_true")
    ap.add_argument("--shims", action="store_true")
    args = ap.parse_args()
    utils = load_reference_utils()
    if args.deep_hub:
        deep_hub_fixture(utils)
        return
    if args.shims:
        shims_fixture(utils)
        return
    if args.betweenness or args.eigenvector:
        if args.betweenness:
            betweenness_fixture(utils)
        if args.eigenvector:
            eigenvector_fixture(utils)
        return
    out = {}

    # 1. micro graphs through utils.shortest_path_length (utils.py:64-81)
    for name, (n, edges, anchors) in micro_graphs().items():
        out[f"micro/{name}/n"] = np.int64(n)
        out[f"micro/{name}/edges"] = np.asarray(edges, dtype=np.int64).reshape(-1, 2).T
        out[f"micro/{name}/anchors"] = np.asarray(anchors, dtype=np.int64)
        out[f"micro/{name}/rows_f64"] = ref_rows(utils, n, edges, anchors)

    # 2. full geodesic pipeline, asymmetric multigraph, seed-42 stochastic anchors
    n, k = 300, 8
    ei = synth.random_digraph(n, 900, seed=7)
    x = np.random.default_rng(11).standard_normal((n, 5)).astype(np.float32)
    data = RefData(torch.tensor(ei), n, torch.tensor(x))
    np.random.seed(42)
    feats = utils.attach_distance_embedding(data, "toy", k, "stochastic", None, 2)
    out["toy/n"] = np.int64(n)
    out["toy/edge_index"] = ei
    out["toy/x"] = x
    out["toy/anchors"] = np.asarray(data.anchor_nodes, dtype=np.int64)
    out["toy/features"] = feats.numpy()
    assert feats.dtype == torch.float32 and tuple(feats.shape) == (n, 5 + k)

    # 2b. same graph, symmetric variant (pull == push direction) with K=70 (two lane words)
    ei_sym = np.concatenate([ei, ei[::-1]], axis=1)
    data = RefData(torch.tensor(ei_sym), n, torch.tensor(x))
    np.random.seed(43)
    data.anchor_nodes = utils.sample_anchor_nodes(data, 70, "stochastic")
    emb = utils.get_geodesic_distance_vector(data, 2)
    out["toysym/edge_index"] = ei_sym
    out["toysym/anchors"] = np.asarray(data.anchor_nodes, dtype=np.int64)
    out["toysym/embedding"] = emb.numpy().astype(np.float32)

    # 3. samplers on a 2000-node graph (ties at the cut are the norm)
    n3 = 2000
    ei3 = synth.chung_lu_symmetric(n3, 12000, 2.2, seed=21)
    ei3 = np.concatenate([ei3, synth.random_digraph(n3, 500, seed=22)], axis=1)  # asymmetric part
    data3 = RefData(torch.tensor(ei3), n3)
    out["samplers/n"] = np.int64(n3)
    out["samplers/edge_index"] = ei3
    for k3 in (1, 16, 64, 256):
        out[f"samplers/degree_centrality/{k3}"] = np.asarray(
            utils.sample_anchor_nodes(data3, k3, "degree_centrality"), dtype=np.int64)
        out[f"samplers/pagerank/{k3}"] = np.asarray(
            utils.sample_anchor_nodes(data3, k3, "pagerank"), dtype=np.int64)
    import networkx as nx
    G3 = utils.to_networkx(data3)
    pr = nx.pagerank(G3)
    out["samplers/pagerank_scores"] = np.asarray([pr[i] for i in range(n3)], dtype=np.float64)
    out["samplers/degree"] = np.asarray([G3.degree(i) for i in range(n3)], dtype=np.int64)
    for k3 in (1, 16, 64, 256):
        out[f"samplers/closeness_centrality/{k3}"] = np.asarray(
            utils.sample_anchor_nodes(data3, k3, "closeness_centrality"), dtype=np.int64)
    for k3 in (1, 16, 64, 256):
        out[f"samplers/clustering_coefficient/{k3}"] = np.asarray(
            utils.sample_anchor_nodes(data3, k3, "clustering_coefficient"), dtype=np.int64)
    cl = nx.clustering(G3)
    out["samplers/clustering_scores"] = np.asarray([cl[i] for i in range(n3)], dtype=np.float64)
    cc = nx.closeness_centrality(G3)
    out["samplers/closeness_scores"] = np.asarray([cc[i] for i in range(n3)], dtype=np.float64)
    np.random.seed(42)
    out["samplers/stochastic_89250_256"] = np.asarray(
        utils.sample_anchor_nodes(RefData(None, 89250), 256, "stochastic"), dtype=np.int64)
    np.random.seed(42)
    out["samplers/stochastic_300_8"] = np.asarray(
        utils.sample_anchor_nodes(RefData(None, 300), 8, "stochastic"), dtype=np.int64)

    # 4. node2vec branch (utils.py:149-180) with torch.load patched to serve the table
    n4, d4, k4 = 400, 128, 12
    table = synth.node2vec_table(n4, d4, seed=3)
    x4 = np.random.default_rng(12).standard_normal((n4, 3)).astype(np.float32)
    real_load = torch.load
    torch.load = lambda *a, **kw: torch.tensor(table)
    try:
        for fn in ("distance", "similarity", "euclidean"):
            data4 = RefData(None, n4, torch.tensor(x4))
            np.random.seed(5)
            f4 = utils.attach_node2vec(data4, "toy", k4, "stochastic", fn, 2)
            out[f"node2vec/{fn}"] = f4.numpy()
            assert f4.dtype == torch.float32, f4.dtype
    finally:
        torch.load = real_load
    np.random.seed(5)
    out["node2vec/anchors"] = np.asarray(np.random.choice(np.arange(n4), k4), dtype=np.int64)
    out["node2vec/table_seed"] = np.int64(3)
    out["node2vec/x"] = x4

    np.savez_compressed(os.path.join(HERE, "reference_small.npz"), **out)
    print("wrote reference_small.npz with", len(out), "arrays")

    if args.pubmed:
        # 5. C1: Pubmed-shape, K=256, reference verbatim with num_workers=6
        import time
        shape = synth.PUBMED_SHAPE
        ei5 = synth.make_graph(shape)
        data5 = RefData(torch.tensor(ei5), shape.num_nodes)
        np.random.seed(42)
        data5.anchor_nodes = utils.sample_anchor_nodes(data5, 256, "stochastic")
        t0 = time.time()
        emb5 = utils.get_geodesic_distance_vector(data5, 6).numpy().astype(np.float32)
        dt = time.time() - t0
        rows = np.arange(0, shape.num_nodes, 97)
        np.savez_compressed(
            os.path.join(HERE, "reference_pubmed_shape.npz"),
            anchors=np.asarray(data5.anchor_nodes, dtype=np.int64),
            sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(emb5).tobytes()).digest(), dtype=np.uint8),
            sample_rows=rows, sample=emb5[rows],
            column_sums=emb5.astype(np.float64).sum(axis=0),
            seconds=np.float64(dt), num_workers=np.int64(6))
        print("wrote reference_pubmed_shape.npz; reference took %.1f s" % dt)


if __name__ == "__main__":
    main()
