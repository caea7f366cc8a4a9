// gp_msbfs.cu — multi-source BFS over bit lanes, one persistent cooperative kernel (sm_100a).
//
// Replaces the N*K nx.shortest_path calls of the reference hot loop (utils.py:69-77, fanned
// out by utils.py:92-114): hops(node -> anchor_j) for all nodes and anchors is computed as a
// level-synchronous BFS from all anchors at once, 64 anchors per uint64 lane word.
//
// Formulation.  Paths run node -> anchor (row = source, utils.py:73), so the frontier moves
// against the edge direction.  In PULL form a row u ORs the frontier words of its
// OUT-neighbours: new[u] = (OR_{u->v} frontier[v]) & ~seen[u].  Each row is owned by exactly
// one thread group, so there are no atomics on the lane state and a level needs one grid
// barrier.  Rows are visited through the degree-ordered list built with the CSR
// (gp_csr.cu): hubs get a whole CTA, medium rows a warp, short rows a thread (pair).
//
// Data layout in HBM/L2 (all uint64, node-major so one neighbour gather is one 32-byte
// sector when wb == 4):
//   seen / frontier ping / frontier pong : [batches][N][wb]
//   distance bit planes p = 0..15        : [batches][N][wb], bit j of plane p = bit p of
//                                          hops(node, anchor j); written only when the lane
//                                          is first reached (at level L every plane with a
//                                          set bit in L is OR-ed with the new mask).
// The uint16 / fp32 matrices are never scattered to: the epilogue (gp_epilogue.cu) decodes
// the planes row by row with fully coalesced stores.
#include "gp_msbfs.cuh"

#include <cooperative_groups.h>
#include <new>

namespace {

struct BfsParams {
    int n;
    int batches;
    int num_anchors;
    const int *__restrict__ rowptr;
    const int *__restrict__ col;
    const int *__restrict__ order;
    const int *__restrict__ meta;   // csr meta words (class boundaries)
    const long long *__restrict__ anchors;
    u64 *seen;
    u64 *fr_a;
    u64 *fr_b;
    u64 *planes;
    long long plane_stride;         // words per plane = batches * n * wb
    u64 *live;                      // [3][GP_BFS_MAX_LANE_WORDS]
    u32 *sync_words;
    int *status;
    u64 *counters;
};

template <int VW>
__device__ __forceinline__ void vload(const u64 *p, u64 (&v)[VW])
{
    if constexpr (VW == 2) {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(p);
        v[0] = t.x;
        v[1] = t.y;
    } else {
        v[0] = *p;
    }
}

template <int VW>
__device__ __forceinline__ void vstore(u64 *p, const u64 (&v)[VW])
{
    if constexpr (VW == 2) {
        *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(v[0], v[1]);
    } else {
        *p = v[0];
    }
}

template <int VW>
struct LevelCtx {
    const u64 *cur;
    u64 *nxt;
    u64 *seen;
    u64 *planes;
    long long plane_stride;
    int level;
    int zplane;  // plane to clear during this sweep (-1: none)
};

// Owner-side update of one row's VW lane words.
template <int VW>
__device__ __forceinline__ void finalize_row(const LevelCtx<VW> &c, size_t off, const u64 (&acc)[VW],
                                             u64 (&seenv)[VW], u64 (&live_acc)[VW])
{
    u64 nw[VW];
    bool any = false;
#pragma unroll
    for (int i = 0; i < VW; ++i) {
        nw[i] = acc[i] & ~seenv[i];
        any |= nw[i] != 0;
    }
    vstore<VW>(c.nxt + off, nw);
    if (c.zplane >= 0) {
        u64 z[VW];
#pragma unroll
        for (int i = 0; i < VW; ++i) z[i] = 0;
        vstore<VW>(c.planes + (size_t)c.zplane * c.plane_stride + off, z);
    }
    if (any) {
#pragma unroll
        for (int i = 0; i < VW; ++i) {
            seenv[i] |= nw[i];
            live_acc[i] |= nw[i];
        }
        vstore<VW>(c.seen + off, seenv);
        for (int lb = c.level; lb; lb &= lb - 1) {
            u64 *pp = c.planes + (size_t)(__ffs(lb) - 1) * c.plane_stride + off;
            u64 t[VW];
            vload<VW>(pp, t);
#pragma unroll
            for (int i = 0; i < VW; ++i) t[i] |= nw[i];
            vstore<VW>(pp, t);
        }
    }
}

// WB lane words per node row; TPE threads share one edge (each loads VW = WB/TPE words).
template <int WB>
__global__ void __launch_bounds__(GP_BFS_THREADS, 2) msbfs_kernel(BfsParams p)
{
    constexpr int TPE = (WB == 4) ? 2 : 1;
    constexpr int VW = WB / TPE;
    constexpr int PAIRS_PER_WARP = 32 / TPE;
    constexpr int PAIRS_PER_CTA = GP_BFS_THREADS / TPE;
    constexpr int WARPS = GP_BFS_THREADS / 32;

    __shared__ u32 s_live32[GP_BFS_MAX_LANE_WORDS * 2];
    __shared__ u64 s_red[WARPS][WB];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = (TPE == 2) ? (lane & 1) : 0;
    const int woff = half * VW;
    const int pairlane = lane / TPE;
    const long long gthreads = (long long)gridDim.x * GP_BFS_THREADS;
    const long long gtid = (long long)blockIdx.x * GP_BFS_THREADS + tid;
    const int gwarp = blockIdx.x * WARPS + warp, total_warps = gridDim.x * WARPS;
    const long long gpair = gtid / TPE, total_pairs = gthreads / TPE;
    const int n = p.n, lw = p.batches * WB;
    u32 bar_target = 0;
    u64 gathers = 0;

    for (int i = tid; i < GP_BFS_MAX_LANE_WORDS * 2; i += GP_BFS_THREADS) s_live32[i] = 0;

    // ---- level 0: seed the anchors (duplicates simply set their own lane bits)
    for (long long j = gtid; j < p.num_anchors; j += gthreads) {
        const long long a = p.anchors[j];
        if (a < 0 || a >= n) {
            atomicOr(&p.status[GP_BFS_ST_ERROR], GP_DEV_ERR_ANCHOR_RANGE);
            continue;
        }
        const int b = (int)(j / (64 * WB)), w = (int)((j / 64) % WB);
        const u64 bit = 1ull << (j & 63);
        const size_t off = ((size_t)b * n + (size_t)a) * WB + w;
        atomicOr(p.seen + off, bit);
        atomicOr(p.fr_a + off, bit);
        atomicOr(p.live + 1 * GP_BFS_MAX_LANE_WORDS + b * WB + w, bit);
    }
    grid_barrier(p.sync_words, bar_target, gridDim.x);

    const int n_large = p.meta[GP_META_N_LARGE];
    const int n_lm = p.meta[GP_META_N_LARGE_MED];
    const int n_med = n_lm - n_large, n_small = n - n_lm;

    int level = 1, max_level = 0;
    while (true) {
        if (level >= (int)GP_UNREACHABLE_U16) {
            if (gtid == 0) atomicOr(&p.status[GP_BFS_ST_ERROR], GP_DEV_ERR_LEVEL_OVERFLOW);
            break;
        }
        LevelCtx<VW> c;
        c.cur = (level & 1) ? p.fr_a : p.fr_b;
        c.nxt = (level & 1) ? p.fr_b : p.fr_a;
        c.seen = p.seen;
        c.planes = p.planes;
        c.plane_stride = p.plane_stride;
        c.level = level;
        c.zplane = ((level + 1) & level) == 0 ? (31 - __clz(level + 1)) : -1;
        if (c.zplane >= GP_BFS_PLANES) c.zplane = -1;
        const u64 *live_r = p.live + (level % 3) * GP_BFS_MAX_LANE_WORDS;
        u64 *live_w = p.live + ((level + 1) % 3) * GP_BFS_MAX_LANE_WORDS;
        u64 *live_z = p.live + ((level + 2) % 3) * GP_BFS_MAX_LANE_WORDS;
        if (blockIdx.x == 0 && tid < lw) live_z[tid] = 0;

        for (int b = 0; b < p.batches; ++b) {
            u64 lv[VW], live_acc[VW];
#pragma unroll
            for (int i = 0; i < VW; ++i) {
                lv[i] = live_r[b * WB + woff + i];
                live_acc[i] = 0;
            }
            const size_t bbase = (size_t)b * n;

            // ---- hubs: one CTA per row
            for (int k = blockIdx.x; k < n_large; k += gridDim.x) {
                const int u = p.order[k];
                const int s = __ldg(p.rowptr + u), e = __ldg(p.rowptr + u + 1);
                const size_t off = (bbase + u) * WB + woff;
                u64 seenv[VW], acc[VW];
                vload<VW>(p.seen + off, seenv);
                bool need = false;
#pragma unroll
                for (int i = 0; i < VW; ++i) {
                    acc[i] = 0;
                    need |= (~seenv[i] & lv[i]) != 0;
                }
                if (need) {
                    for (int j = s + tid / TPE; j < e; j += PAIRS_PER_CTA) {
                        const int v = __ldg(p.col + j);
                        u64 t[VW];
                        vload<VW>(c.cur + (bbase + v) * WB + woff, t);
#pragma unroll
                        for (int i = 0; i < VW; ++i) acc[i] |= t[i];
                        gathers += VW;
                    }
                }
#pragma unroll
                for (int m = TPE; m < 32; m <<= 1)
#pragma unroll
                    for (int i = 0; i < VW; ++i) acc[i] |= shfl_xor_u64(acc[i], m);
                if (lane < TPE)
#pragma unroll
                    for (int i = 0; i < VW; ++i) s_red[warp][woff + i] = acc[i];
                __syncthreads();
                if (warp == 0 && lane < TPE) {
#pragma unroll
                    for (int w = 1; w < WARPS; ++w)
#pragma unroll
                        for (int i = 0; i < VW; ++i) acc[i] |= s_red[w][woff + i];
                    finalize_row<VW>(c, off, acc, seenv, live_acc);
                }
                __syncthreads();
            }

            // ---- medium rows: one warp per row
            for (int k = gwarp; k < n_med; k += total_warps) {
                const int u = p.order[n_large + k];
                const int s = __ldg(p.rowptr + u), e = __ldg(p.rowptr + u + 1);
                const size_t off = (bbase + u) * WB + woff;
                u64 seenv[VW], acc[VW];
                vload<VW>(p.seen + off, seenv);
                bool need = false;
#pragma unroll
                for (int i = 0; i < VW; ++i) {
                    acc[i] = 0;
                    need |= (~seenv[i] & lv[i]) != 0;
                }
                if (need) {
                    for (int j = s + pairlane; j < e; j += PAIRS_PER_WARP) {
                        const int v = __ldg(p.col + j);
                        u64 t[VW];
                        vload<VW>(c.cur + (bbase + v) * WB + woff, t);
#pragma unroll
                        for (int i = 0; i < VW; ++i) acc[i] |= t[i];
                        gathers += VW;
                    }
                }
#pragma unroll
                for (int m = TPE; m < 32; m <<= 1)
#pragma unroll
                    for (int i = 0; i < VW; ++i) acc[i] |= shfl_xor_u64(acc[i], m);
                if (lane < TPE) finalize_row<VW>(c, off, acc, seenv, live_acc);
            }

            // ---- short rows: one thread (pair) per row, no cross-lane traffic
            for (long long k = gpair; k < n_small; k += total_pairs) {
                const int u = p.order[n_lm + k];
                const int s = __ldg(p.rowptr + u), e = __ldg(p.rowptr + u + 1);
                const size_t off = (bbase + u) * WB + woff;
                u64 seenv[VW], acc[VW];
                vload<VW>(p.seen + off, seenv);
                bool need = false;
#pragma unroll
                for (int i = 0; i < VW; ++i) {
                    acc[i] = 0;
                    need |= (~seenv[i] & lv[i]) != 0;
                }
                if (need) {
#pragma unroll 4
                    for (int j = s; j < e; ++j) {
                        const int v = __ldg(p.col + j);
                        u64 t[VW];
                        vload<VW>(c.cur + (bbase + v) * WB + woff, t);
#pragma unroll
                        for (int i = 0; i < VW; ++i) acc[i] |= t[i];
                    }
                    gathers += (u64)VW * (u64)(e - s);
                }
                finalize_row<VW>(c, off, acc, seenv, live_acc);
            }

            // ---- fold this batch's newly reached lanes into the CTA's live words
#pragma unroll
            for (int i = 0; i < VW; ++i) {
                u64 v = live_acc[i];
#pragma unroll
                for (int m = TPE; m < 32; m <<= 1) v |= shfl_xor_u64(v, m);
                if (lane < TPE && v) {
                    atomicOr(&s_live32[(b * WB + woff + i) * 2], (u32)v);
                    atomicOr(&s_live32[(b * WB + woff + i) * 2 + 1], (u32)(v >> 32));
                }
            }
        }
        __syncthreads();
        if (tid < lw) {
            const u64 v = ((u64)s_live32[tid * 2 + 1] << 32) | s_live32[tid * 2];
            if (v) atomicOr(live_w + tid, v);
            s_live32[tid * 2] = 0;
            s_live32[tid * 2 + 1] = 0;
        }
        grid_barrier(p.sync_words, bar_target, gridDim.x);
        const int any = __syncthreads_or(tid < lw && live_w[tid] != 0);
        if (!any) break;
        max_level = level;
        ++level;
    }

    // ---- stats
    for (int m = 16; m; m >>= 1) gathers += shfl_xor_u64(gathers, m);
    if (lane == 0 && gathers) atomicAdd(p.counters, gathers);
    if (gtid == 0) {
        p.status[GP_BFS_ST_MAX_LEVEL] = max_level;
        p.status[GP_BFS_ST_LEVELS] = level < (int)GP_UNREACHABLE_U16 ? level : level - 1;
        p.status[GP_BFS_ST_PULL] = level < (int)GP_UNREACHABLE_U16 ? level : level - 1;
    }
}

template <int WB>
int launch_bfs(gp_msbfs *h, const BfsParams &p, cudaStream_t stream)
{
    if (h->grid_blocks == 0) {
        int occ = 0;
        GP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, msbfs_kernel<WB>, GP_BFS_THREADS, 0));
        GP_REQUIRE(occ >= 1, GP_ERR_CUDA, "msbfs kernel does not fit on an SM");
        if (const char *s = getenv("GP_BFS_CTAS_PER_SM")) {
            const int want = atoi(s);
            if (want >= 1 && want < occ) occ = want;
        }
        h->grid_blocks = occ * gp_sm_count();
    }
    BfsParams pp = p;
    void *args[] = {&pp};
    gp_count_launch();
    GP_CUDA_CHECK(cudaEventRecord(h->ev_start, stream));
    GP_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)msbfs_kernel<WB>, dim3(h->grid_blocks),
                                              dim3(GP_BFS_THREADS), args, 0, stream));
    GP_CUDA_CHECK(cudaEventRecord(h->ev_stop, stream));
    return GP_OK;
}

}  // namespace

static void bfs_config(int64_t k, int *wb, int *batches)
{
    *wb = k <= 64 ? 1 : (k <= 128 ? 2 : 4);
    *batches = (int)gp_ceil_div(k > 0 ? k : 1, 64 * (*wb));
}

extern "C" int gp_msbfs_create(const gp_csr_t *csr, int64_t max_anchors, gp_msbfs_t **out)
{
    GP_REQUIRE(out != nullptr, GP_ERR_INVALID, "gp_msbfs_create: out is NULL");
    *out = nullptr;
    GP_REQUIRE(csr != nullptr, GP_ERR_INVALID, "gp_msbfs_create: csr is NULL");
    GP_REQUIRE(max_anchors >= 0, GP_ERR_INVALID, "gp_msbfs_create: negative anchor count");
    int wb, batches;
    bfs_config(max_anchors, &wb, &batches);
    GP_REQUIRE((int64_t)wb * batches <= GP_BFS_MAX_LANE_WORDS, GP_ERR_UNSUPPORTED,
               "gp_msbfs_create: %lld anchors exceed the %d-anchor limit per GPU", (long long)max_anchors,
               GP_BFS_MAX_LANE_WORDS * 64);
    gp_msbfs *h = new (std::nothrow) gp_msbfs();
    GP_REQUIRE(h != nullptr, GP_ERR_OOM, "gp_msbfs_create: host allocation failed");
    h->csr = csr;
    h->num_nodes = csr->num_nodes;
    h->max_anchors = max_anchors;
    h->cap_words_per_node = (int64_t)wb * batches;
    const size_t words = (size_t)h->cap_words_per_node * (size_t)(h->num_nodes > 0 ? h->num_nodes : 1);
    int rc = GP_OK;
    auto alloc = [&](void **ptr, size_t bytes) {
        if (rc != GP_OK) return;
        cudaError_t e = cudaMalloc(ptr, bytes > 0 ? bytes : 16);
        if (e != cudaSuccess) {
            gp_set_error("gp_msbfs_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            rc = (e == cudaErrorMemoryAllocation) ? GP_ERR_OOM : GP_ERR_CUDA;
        }
    };
    // seen is plane 0 of the result; the 16 distance planes follow it contiguously
    alloc((void **)&h->seen, (size_t)(1 + GP_BFS_PLANES) * words * sizeof(u64));
    alloc((void **)&h->fr_a, 2 * words * sizeof(u64));
    alloc((void **)&h->live, 3 * GP_BFS_MAX_LANE_WORDS * sizeof(u64));
    alloc((void **)&h->queue, (size_t)(h->num_nodes + 1) * sizeof(int));
    alloc((void **)&h->sync_words, 64 * sizeof(u32));
    alloc((void **)&h->status, GP_BFS_ST_WORDS * sizeof(int));
    alloc((void **)&h->counters, 4 * sizeof(u64));
    if (rc == GP_OK && (cudaEventCreate(&h->ev_start) != cudaSuccess || cudaEventCreate(&h->ev_stop) != cudaSuccess)) {
        gp_set_error("gp_msbfs_create: cudaEventCreate failed");
        rc = GP_ERR_CUDA;
    }
    if (rc != GP_OK) {
        gp_msbfs_free(h);
        return rc;
    }
    *out = h;
    return GP_OK;
}

extern "C" int gp_msbfs_free(gp_msbfs_t *h)
{
    if (!h) return GP_OK;
    cudaFree(h->seen);
    cudaFree(h->fr_a);
    cudaFree(h->live);
    cudaFree(h->queue);
    cudaFree(h->sync_words);
    cudaFree(h->status);
    cudaFree(h->counters);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_stop) cudaEventDestroy(h->ev_stop);
    delete h;
    return GP_OK;
}

extern "C" int gp_msbfs_run(gp_msbfs_t *h, const int64_t *d_anchors, int64_t num_anchors, gp_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(h != nullptr, GP_ERR_INVALID, "gp_msbfs_run: handle is NULL");
    GP_REQUIRE(h->csr->built, GP_ERR_INVALID, "gp_msbfs_run: the CSR has not been built");
    GP_REQUIRE(h->csr->num_nodes == h->num_nodes, GP_ERR_INVALID, "gp_msbfs_run: CSR changed size");
    GP_REQUIRE(num_anchors >= 0 && num_anchors <= h->max_anchors, GP_ERR_INVALID,
               "gp_msbfs_run: %lld anchors exceed max_anchors %lld", (long long)num_anchors,
               (long long)h->max_anchors);
    GP_REQUIRE(num_anchors == 0 || d_anchors != nullptr, GP_ERR_INVALID, "gp_msbfs_run: anchors is NULL");
    int wb, batches;
    bfs_config(num_anchors, &wb, &batches);
    h->wb = wb;
    h->batches = batches;
    h->num_anchors = num_anchors;
    h->ran = false;
    const int64_t n = h->num_nodes;
    const size_t words = (size_t)wb * batches * (size_t)n;
    h->fr_b = h->fr_a + words;
    h->planes = h->seen + words;
    GP_CUDA_CHECK(cudaMemsetAsync(h->status, 0, GP_BFS_ST_WORDS * sizeof(int), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(h->counters, 0, 4 * sizeof(u64), stream));
    if (n == 0 || num_anchors == 0) {
        h->ran = true;
        return GP_OK;
    }
    // seen + distance plane 0 are contiguous; frontier ping separately
    GP_CUDA_CHECK(cudaMemsetAsync(h->seen, 0, 2 * words * sizeof(u64), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(h->fr_a, 0, words * sizeof(u64), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(h->live, 0, 3 * GP_BFS_MAX_LANE_WORDS * sizeof(u64), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(h->sync_words, 0, 64 * sizeof(u32), stream));
    BfsParams p;
    p.n = (int)n;
    p.batches = batches;
    p.num_anchors = (int)num_anchors;
    p.rowptr = h->csr->rowptr_out;
    p.col = h->csr->col_out;
    p.order = h->csr->order;
    p.meta = h->csr->meta;
    p.anchors = (const long long *)d_anchors;
    p.seen = h->seen;
    p.fr_a = h->fr_a;
    p.fr_b = h->fr_b;
    p.planes = h->planes;
    p.plane_stride = (long long)words;
    p.live = h->live;
    p.sync_words = h->sync_words;
    p.status = h->status;
    p.counters = h->counters;
    int prev_grid = h->grid_blocks;
    static int grid_for_wb[5] = {0, 0, 0, 0, 0};
    h->grid_blocks = grid_for_wb[wb];
    int rc = wb == 1 ? launch_bfs<1>(h, p, stream) : wb == 2 ? launch_bfs<2>(h, p, stream)
                                                             : launch_bfs<4>(h, p, stream);
    grid_for_wb[wb] = h->grid_blocks;
    (void)prev_grid;
    GP_TRY(rc);
    h->ran = true;
    return GP_OK;
}

extern "C" int gp_msbfs_stats(gp_msbfs_t *h, gp_msbfs_stats_t *stats, gp_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(h != nullptr && stats != nullptr, GP_ERR_INVALID, "gp_msbfs_stats: NULL argument");
    GP_REQUIRE(h->ran, GP_ERR_INVALID, "gp_msbfs_stats: gp_msbfs_run has not been called");
    int st[GP_BFS_ST_WORDS];
    u64 cnt[4];
    int csr_err = 0;
    GP_CUDA_CHECK(cudaMemcpyAsync(st, h->status, sizeof(st), cudaMemcpyDeviceToHost, stream));
    GP_CUDA_CHECK(cudaMemcpyAsync(cnt, h->counters, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
    GP_CUDA_CHECK(cudaMemcpyAsync(&csr_err, &h->csr->meta[GP_META_ERROR], sizeof(int), cudaMemcpyDeviceToHost,
                                  stream));
    GP_CUDA_CHECK(cudaStreamSynchronize(stream));
    stats->num_anchors = h->num_anchors;
    stats->lane_words = (int64_t)h->wb * h->batches;
    stats->max_level = st[GP_BFS_ST_MAX_LEVEL];
    stats->levels_run = st[GP_BFS_ST_LEVELS];
    stats->pull_levels = st[GP_BFS_ST_PULL];
    stats->push_levels = st[GP_BFS_ST_PUSH];
    stats->edges_examined = (int64_t)cnt[0] + (int64_t)cnt[1];
    stats->grid_blocks = h->grid_blocks;
    GP_REQUIRE(!(csr_err & GP_DEV_ERR_EDGE_RANGE), GP_ERR_INDEX_RANGE,
               "edge_index holds an entry outside [0, %lld)", (long long)h->num_nodes);
    GP_REQUIRE(!(st[GP_BFS_ST_ERROR] & GP_DEV_ERR_ANCHOR_RANGE), GP_ERR_INDEX_RANGE,
               "anchor index outside [0, %lld)", (long long)h->num_nodes);
    GP_REQUIRE(!(st[GP_BFS_ST_ERROR] & GP_DEV_ERR_LEVEL_OVERFLOW), GP_ERR_LEVEL_OVERFLOW,
               "a hop distance reached 65535 and does not fit the uint16 distance matrix");
    return GP_OK;
}

extern "C" int gp_msbfs_kernel_ms(gp_msbfs_t *h, float *ms)
{
    GP_REQUIRE(h != nullptr && ms != nullptr, GP_ERR_INVALID, "gp_msbfs_kernel_ms: NULL argument");
    GP_REQUIRE(h->ran, GP_ERR_INVALID, "gp_msbfs_kernel_ms: gp_msbfs_run has not been called");
    *ms = 0.0f;
    if (h->num_nodes == 0 || h->num_anchors == 0) return GP_OK;
    GP_CUDA_CHECK(cudaEventSynchronize(h->ev_stop));
    GP_CUDA_CHECK(cudaEventElapsedTime(ms, h->ev_start, h->ev_stop));
    return GP_OK;
}

extern "C" int gp_msbfs_planes(gp_msbfs_t *h, const uint64_t **d_planes, int64_t *plane_stride_words,
                               int32_t *num_planes, int32_t *batches, int32_t *words_per_batch,
                               gp_stream_t stream_)
{
    GP_REQUIRE(h != nullptr && d_planes && plane_stride_words && num_planes && batches && words_per_batch,
               GP_ERR_INVALID, "gp_msbfs_planes: NULL argument");
    gp_msbfs_stats_t st;
    GP_TRY(gp_msbfs_stats(h, &st, stream_));
    int bits = 0;
    while ((1 << bits) <= st.max_level) ++bits;
    *d_planes = (const uint64_t *)h->seen;
    *plane_stride_words = (int64_t)h->wb * h->batches * h->num_nodes;
    *num_planes = 1 + bits;
    *batches = h->batches;
    *words_per_batch = h->wb;
    return GP_OK;
}
