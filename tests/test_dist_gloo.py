"""world_size-2 gloo test of the anchor-sharding host logic (no GPU).

Each rank encodes the oracle's hops for ITS anchor shard in the library's result-block layout
(R[0] reached mask, R[l] first-reached-at-hop-l), the product code agrees on the array count and
all-gathers the blocks, and a numpy restatement of the epilogue's column permutation must
reproduce the oracle's full [N, K] matrix.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphpope_b200 import distributed as gpd
from graphpope_b200 import synth


def encode_block(hops, wb, num_arrays):
    """uint16 [N, Kr] -> int64 [num_arrays, batches*N*wb] in the gp_msbfs result layout."""
    n, kr = hops.shape
    batches = -(-kr // (64 * wb))
    blk = np.zeros((num_arrays, batches, n, wb), dtype=np.uint64)
    for j in range(kr):
        b, w, bit = j // (64 * wb), (j // 64) % wb, np.uint64(j % 64)
        d = hops[:, j]
        reach = d != 0xFFFF
        blk[0, b, reach, w] |= np.uint64(1) << bit
        for l in range(1, num_arrays):
            blk[l, b, d == l, w] |= np.uint64(1) << bit
    return torch.from_numpy(blk.reshape(num_arrays, -1).view(np.int64))


def decode_gathered(g, n, kr, wb):
    """numpy restatement of gp_decode_gathered's addressing: [G, P, words] -> uint16 [N, G*Kr]."""
    world, p, _ = g.shape
    batches = -(-kr // (64 * wb))
    blk = g.numpy().view(np.uint64).reshape(world, p, batches, n, wb)
    out = np.full((n, world * kr), 0xFFFF, dtype=np.uint16)
    for r in range(world):
        for j in range(kr):
            b, w, bit = j // (64 * wb), (j // 64) % wb, np.uint64(j % 64)
            reach = ((blk[r, 0, b, :, w] >> bit) & np.uint64(1)).astype(bool)
            d = np.zeros(n, dtype=np.uint16)
            for l in range(1, p):
                d[((blk[r, l, b, :, w] >> bit) & np.uint64(1)).astype(bool)] = l
            out[reach, r * kr + j] = d[reach]
    return out


def _worker(rank, world, port, k_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cbfs
        n = 400
        ei = synth.random_digraph(n, 1500, seed=3)
        anchors = np.random.default_rng(5).integers(0, n, k_total)
        lo, hi = gpd.shard_bounds(k_total, world, rank)
        hops = cbfs.bfs_hops(cbfs.InCsr(ei, n), anchors[lo:hi])
        local_max = int(hops[hops != 0xFFFF].max())
        wb = 1 if hi - lo <= 64 else (2 if hi - lo <= 128 else 4)
        num = gpd.agree_num_planes(1 + local_max)          # max over ranks
        local = encode_block(hops, wb, num)
        gathered = gpd.gather_planes(local)                 # [G, P, words]
        full = decode_gathered(gathered, n, hi - lo, wb)
        want = cbfs.bfs_hops(cbfs.InCsr(ei, n), anchors)
        ret[rank] = bool(np.array_equal(full, want)) and num >= 1 + local_max
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k_total", [6, 200])
def test_anchor_sharding_world_size_2(k_total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, k_total, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_bounds():
    assert [gpd.shard_bounds(1024, 8, r) for r in (0, 7)] == [(0, 128), (896, 1024)]
    # ragged: ceil(K / G) per rank, the tail ranks hold fewer (the reference accepts any K)
    assert [gpd.shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [gpd.shard_bounds(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    padded, per = gpd.pad_anchors(np.arange(10), 4)
    assert per == 8 and len(padded) == 32 and padded[:10].tolist() == list(range(10)) and set(padded[10:]) == {9}
    padded, per = gpd.pad_anchors(torch.arange(1024), 8)
    assert per == 128 and padded.numel() == 1024


def _shared_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, f, k = 1000, 7, 10
        shared = gpd.SharedHostMatrix(n, f + k, register=False)  # no CUDA here: the mapping alone
        lo, hi = gpd.shard_bounds(k, world, rank)
        r0, r1 = gpd.row_slice(n, world, rank)
        x = torch.arange(n * f, dtype=torch.float32).view(n, f)
        shared.tensor[:, f + lo:f + hi] = float(rank + 1)       # this rank's column block
        shared.tensor[r0:r1, :f] = x[r0:r1]                     # this rank's rows of x
        shared.barrier()
        want = torch.cat([x, torch.cat([torch.full((n, gpd.shard_bounds(k, world, r)[1] - gpd.shard_bounds(k, world, r)[0]),
                                                   float(r + 1)) for r in range(world)], 1)], 1)
        ret[rank] = bool(torch.equal(shared.tensor, want)) and not os.path.exists(shared.path)
        shared.barrier()
        shared.close()
    finally:
        dist.destroy_process_group()


def test_node_shared_host_matrix_world_size_2():
    """SharedHostMatrix: two processes map the same POSIX shared-memory matrix; each writes its column block and its
    row range of x; after the barrier both see the whole [N, F + K] matrix and the name is unlinked."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ret = mp.Manager().dict()
    mp.spawn(_shared_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
