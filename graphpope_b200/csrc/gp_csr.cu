// gp_csr.cu — device CSR builder: the to_networkx(data) replacement (reference utils.py:121,
// also :27,33,39,45,51,57).  DiGraph semantics: nodes 0..N-1, parallel edges collapse,
// self-loops stay, no symmetrisation unless GP_CSR_SYMMETRIZE.
//
//   edge_index int64 [2,E] --pack--> (src << nb | dst) keys --onesweep radix sort-->
//   --fused unique--> sorted distinct keys --unpack--> col_out + rowptr_out (lower bounds)
//   --degree keys + 2-pass sort--> degree-ordered row list with class boundaries.
// The in-edge CSR (push direction, PageRank pull) is a stable re-sort of the distinct
// keys by dst only, built on demand.  Everything is stream-ordered; counts stay on device.
#include "gp_internal.h"

#include <new>

namespace {

__global__ void pack_keys_kernel(const long long *__restrict__ ei, long long e, long long n, int nb,
                                 int symmetrize, u64 *__restrict__ keys, int *meta)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += stride) {
        long long s = ei[i], d = ei[e + i];
        if (s < 0 || d < 0 || s >= n || d >= n) {
            bad = true;
            s = 0;
            d = 0;
        }
        keys[i] = ((u64)s << nb) | (u64)d;
        if (symmetrize) keys[e + i] = ((u64)d << nb) | (u64)s;
    }
    if (__any_sync(FULL_MASK, bad) && lane_id() == 0) atomicOr(&meta[GP_META_ERROR], GP_DEV_ERR_EDGE_RANGE);
}

__device__ __forceinline__ u32 lower_bound_u64(const u64 *__restrict__ a, u32 n, u64 key)
{
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// col[i] = low half of key i; rowptr[r] = first key whose high half is >= r.
__global__ void unpack_csr_kernel(const u64 *__restrict__ ukeys, const int *__restrict__ meta, long long n,
                                  int nb, long long cap, int *__restrict__ rowptr, int *__restrict__ col)
{
    const u32 m = (u32)meta[GP_META_NUM_EDGES];
    const u64 mask = (1ull << nb) - 1ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long work = cap > n + 1 ? cap : n + 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < work; i += stride) {
        if (i < (long long)m) col[i] = (int)(ukeys[i] & mask);
        if (i <= n) rowptr[i] = (int)lower_bound_u64(ukeys, m, (u64)i << nb);
    }
}

__global__ void degree_keys_kernel(const int *__restrict__ rowptr, long long n, u64 *__restrict__ okeys,
                                   int *meta)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    int mx = 0;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += stride) {
        const int deg = rowptr[u + 1] - rowptr[u];
        mx = max(mx, deg);
        const u32 capped = deg > 65535 ? 65535u : (u32)deg;
        okeys[u] = ((u64)(65535u - capped) << 32) | (u64)(u32)u;  // ascending key = descending degree
    }
    mx = __reduce_max_sync(FULL_MASK, mx);
    if (lane_id() == 0 && mx > 0) atomicMax(&meta[GP_META_MAX_DEGREE], mx);
}

__global__ void order_kernel(const u64 *__restrict__ okeys, long long n, int *__restrict__ order, int *meta)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = gid; i < n; i += stride) order[i] = (int)(u32)okeys[i];
    if (gid < GP_NUM_CLASSES) {
        // rank boundary: number of rows with degree > thr  <=>  key high half < 65535 - thr
        int cnt = (int)n;
        if (gid < GP_NUM_CLASSES - 1) {
            const int thr = GP_CHUNK_EDGES >> gid;  // 128, 64, 32, 16, 8, 4
            cnt = (int)lower_bound_u64(okeys, (u32)n, (u64)(65535u - (u32)thr) << 32);
        }
        meta[GP_META_RANK + gid] = cnt;
        if (gid == 0) meta[GP_META_NUM_HUB_ROWS] = cnt;
    }
}

// Exclusive prefix over the hub rows (degree order) of their chunk counts; one block, running carry.
__global__ void __launch_bounds__(1024)
hub_scan_kernel(const int *__restrict__ order, const int *__restrict__ rowptr, int *meta,
                int *__restrict__ chunk_off)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int nh = meta[GP_META_NUM_HUB_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nh; base += 1024) {
        const int k = base + tid;
        int v = 0;
        if (k < nh) {
            const int u = order[k];
            v = (rowptr[u + 1] - rowptr[u] + GP_CHUNK_EDGES - 1) / GP_CHUNK_EDGES;
        }
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL_MASK, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        int wbase = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int t = s_warp[w];
            if (w < warp) wbase += t;
            tot += t;
        }
        const int carry = s_carry;
        if (k < nh) chunk_off[k] = carry + wbase + x - v;
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (tid == 0) {
        const int chunks = s_carry;
        chunk_off[nh] = chunks;
        // class c: descriptors [ent_base[c], ent_base[c+1]), G(c) slots each, regions aligned
        int ent = 0, slot = 0;
        for (int c = 0; c < GP_NUM_CLASSES; ++c) {
            const int rows = c == 0 ? chunks : meta[GP_META_RANK + c] - meta[GP_META_RANK + c - 1];
            const int g = c <= 1 ? 32 : (32 >> (c - 1));
            meta[GP_META_ENT_BASE + c] = ent;
            meta[GP_META_SLOT_BASE + c] = slot;
            ent += rows;
            slot += (rows * g + GP_SLOT_ALIGN - 1) / GP_SLOT_ALIGN * GP_SLOT_ALIGN;
        }
        meta[GP_META_ENT_BASE + GP_NUM_CLASSES] = ent;
        meta[GP_META_SLOT_BASE + GP_NUM_CLASSES] = slot;
    }
}

// One descriptor per row (degree order), hub rows expanded into GP_CHUNK_EDGES-edge chunks.
__global__ void desc_kernel(const int *__restrict__ order, const int *__restrict__ rowptr,
                            const int *__restrict__ meta, const int *__restrict__ chunk_off, long long n,
                            int4 *__restrict__ desc)
{
    const int nh = meta[GP_META_NUM_HUB_ROWS];
    const int chunks = chunk_off[nh];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const int u = order[k];
        const int s = rowptr[u], d = rowptr[u + 1] - s;
        if (k >= nh) {
            desc[chunks + (k - nh)] = make_int4(u, s, d, -1);
        } else {
            const int nch = (d + GP_CHUNK_EDGES - 1) / GP_CHUNK_EDGES;
            const int off = chunk_off[k];
            for (int c = 0; c < nch; ++c) {
                const int cnt = min(GP_CHUNK_EDGES, d - c * GP_CHUNK_EDGES);
                desc[off + c] = make_int4(u, s + c * GP_CHUNK_EDGES, cnt | (nch << 8) | (c == 0 ? 1 << 30 : 0), (int)k);
            }
        }
    }
}

__global__ void transpose_keys_kernel(const u64 *__restrict__ ukeys, const int *__restrict__ meta, int nb,
                                      long long cap, u64 *__restrict__ tkeys)
{
    const u32 m = (u32)meta[GP_META_NUM_EDGES];
    const u64 mask = (1ull << nb) - 1ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride)
        if (i < (long long)m) {
            const u64 k = ukeys[i];
            tkeys[i] = ((k & mask) << nb) | (k >> nb);
        }
}

__global__ void symmetric_check_kernel(const u64 *__restrict__ ukeys, const u64 *__restrict__ tkeys,
                                       long long cap, int *meta)
{
    const u32 m = (u32)meta[GP_META_NUM_EDGES];
    const long long stride = (long long)gridDim.x * blockDim.x;
    bool diff = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride)
        if (i < (long long)m && ukeys[i] != tkeys[i]) diff = true;
    if (__any_sync(FULL_MASK, diff) && lane_id() == 0) atomicExch(&meta[GP_META_IS_SYMMETRIC], 0);
}

__global__ void set_meta_kernel(int *meta, int idx, int value) { meta[idx] = value; }

int launch_blocks(int64_t work, int threads)
{
    int64_t b = gp_ceil_div(work > 0 ? work : 1, threads);
    const int64_t cap = (int64_t)gp_sm_count() * 8;
    return (int)(b < cap ? b : cap);
}

}  // namespace

extern "C" int gp_csr_create(int64_t num_nodes, int64_t edge_capacity, uint32_t flags, gp_csr_t **out)
{
    GP_REQUIRE(out != nullptr, GP_ERR_INVALID, "gp_csr_create: out is NULL");
    *out = nullptr;
    GP_REQUIRE(num_nodes >= 0 && edge_capacity >= 0, GP_ERR_INVALID, "gp_csr_create: negative size");
    GP_REQUIRE(num_nodes < (1ll << 31) - 1, GP_ERR_UNSUPPORTED, "gp_csr_create: num_nodes must fit int32");
    const int64_t kcap = edge_capacity * ((flags & GP_CSR_SYMMETRIZE) ? 2 : 1);
    GP_REQUIRE(kcap < (1ll << 30) - 1, GP_ERR_UNSUPPORTED,
               "gp_csr_create: %lld edge keys exceed the 2^30-1 limit of this build", (long long)kcap);
    gp_csr *c = new (std::nothrow) gp_csr();
    GP_REQUIRE(c != nullptr, GP_ERR_OOM, "gp_csr_create: host allocation failed");
    c->num_nodes = num_nodes;
    c->edge_capacity = edge_capacity;
    c->key_capacity = kcap;
    c->flags = flags;
    int nb = 1;
    while ((1ll << nb) < num_nodes) ++nb;
    c->node_bits = nb;
    c->hub_capacity = kcap / (GP_CHUNK_EDGES + 1) + 1;
    c->desc_capacity = num_nodes + kcap / GP_CHUNK_EDGES + 2;
    const size_t kc = (size_t)(kcap > 0 ? kcap : 1), nn = (size_t)num_nodes;
    int rc = GP_OK;
    auto alloc = [&](void **p, size_t bytes) {
        if (rc != GP_OK) return;
        cudaError_t e = cudaMalloc(p, bytes > 0 ? bytes : 16);
        if (e != cudaSuccess) {
            gp_set_error("gp_csr_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            rc = (e == cudaErrorMemoryAllocation) ? GP_ERR_OOM : GP_ERR_CUDA;
        }
    };
    alloc((void **)&c->keys, kc * sizeof(u64));
    alloc((void **)&c->ukeys, kc * sizeof(u64));
    alloc((void **)&c->rowptr_out, (nn + 1) * sizeof(int));
    alloc((void **)&c->col_out, kc * sizeof(int));
    alloc((void **)&c->rowptr_in, (nn + 1) * sizeof(int));
    alloc((void **)&c->col_in, kc * sizeof(int));
    alloc((void **)&c->order, (nn + 1) * sizeof(int));
    alloc((void **)&c->okeys, (nn + 1) * sizeof(u64));
    alloc((void **)&c->desc, (size_t)c->desc_capacity * sizeof(int4));
    alloc((void **)&c->hub_chunk_off, (size_t)(c->hub_capacity + 1) * sizeof(int));
    alloc((void **)&c->meta, GP_META_WORDS * sizeof(int));
    alloc((void **)&c->uniq_status, (size_t)(gp_ceil_div((int64_t)kc, GP_SORT_TILE) + 2) * sizeof(u32));
    if (rc == GP_OK) rc = gp_sort_workspace_create(&c->sort_ws, (int64_t)(kc > nn ? kc : nn), false);
    if (rc != GP_OK) {
        gp_csr_free(c);
        return rc;
    }
    *out = c;
    return GP_OK;
}

extern "C" int gp_csr_free(gp_csr_t *c)
{
    if (!c) return GP_OK;
    gp_drop_graphs(c);
    cudaFree(c->keys);
    cudaFree(c->ukeys);
    cudaFree(c->rowptr_out);
    cudaFree(c->col_out);
    cudaFree(c->rowptr_in);
    cudaFree(c->col_in);
    cudaFree(c->order);
    cudaFree(c->okeys);
    cudaFree(c->desc);
    cudaFree(c->hub_chunk_off);
    cudaFree(c->meta);
    cudaFree(c->uniq_status);
    gp_sort_workspace_free(&c->sort_ws);
    delete c;
    return GP_OK;
}

extern "C" int gp_csr_build(gp_csr_t *c, const int64_t *d_edge_index, int64_t num_edges, gp_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(c != nullptr, GP_ERR_INVALID, "gp_csr_build: csr is NULL");
    GP_REQUIRE(num_edges >= 0 && num_edges <= c->edge_capacity, GP_ERR_INVALID,
               "gp_csr_build: %lld edges exceed the capacity %lld given to gp_csr_create",
               (long long)num_edges, (long long)c->edge_capacity);
    GP_REQUIRE(num_edges == 0 || d_edge_index != nullptr, GP_ERR_INVALID, "gp_csr_build: edge_index is NULL");
    const int sym = (c->flags & GP_CSR_SYMMETRIZE) ? 1 : 0;
    const int64_t nkeys = num_edges * (sym ? 2 : 1);
    const int nb = c->node_bits;
    const int64_t n = c->num_nodes;
    c->num_input_edges = num_edges;
    c->built = false;
    c->in_built = false;
    GP_CUDA_CHECK(cudaMemsetAsync(c->meta, 0, GP_META_WORDS * sizeof(int), stream));
    if (num_edges > 0)
        GP_LAUNCH(pack_keys_kernel, launch_blocks(num_edges, 256), 256, 0, stream, (const long long *)d_edge_index, num_edges, n, nb, sym, c->keys, c->meta);
    u64 *sorted = c->keys;
    GP_TRY(gp_radix_sort(&c->sort_ws, c->keys, nullptr, nullptr, nkeys, 0, 2 * nb, stream, &sorted, nullptr));
    GP_TRY(gp_unique_sorted(sorted, c->ukeys, nullptr, nkeys, (u32 *)&c->meta[GP_META_NUM_EDGES],
                            c->uniq_status, stream));
    const int64_t work = nkeys > n + 1 ? nkeys : n + 1;
    GP_LAUNCH(unpack_csr_kernel, launch_blocks(work, 256), 256, 0, stream, c->ukeys, c->meta, n, nb, nkeys,
                                                                    c->rowptr_out, c->col_out);
    if (n > 0) {
        GP_LAUNCH(degree_keys_kernel, launch_blocks(n, 256), 256, 0, stream, c->rowptr_out, n, c->okeys, c->meta);
        u64 *osorted = c->okeys;
        GP_TRY(gp_radix_sort(&c->sort_ws, c->okeys, nullptr, nullptr, n, 32, 48, stream, &osorted, nullptr));
        GP_LAUNCH(order_kernel, launch_blocks(n, 256), 256, 0, stream, osorted, n, c->order, c->meta);
        GP_LAUNCH(hub_scan_kernel, 1, 1024, 0, stream, c->order, c->rowptr_out, c->meta, c->hub_chunk_off);
        GP_LAUNCH(desc_kernel, launch_blocks(n, 256), 256, 0, stream, c->order, c->rowptr_out, c->meta,
                  c->hub_chunk_off, n, c->desc);
    }
    GP_CUDA_CHECK(cudaGetLastError());
    c->built = true;
    return GP_OK;
}

int gp_csr_ensure_in(gp_csr *c, cudaStream_t stream)
{
    GP_REQUIRE(c != nullptr && c->built, GP_ERR_INVALID, "in-edge CSR requested before gp_csr_build");
    if (c->in_built) return GP_OK;
    const int nb = c->node_bits;
    const int64_t n = c->num_nodes;
    const int64_t cap = c->num_input_edges * ((c->flags & GP_CSR_SYMMETRIZE) ? 2 : 1);
    // c->keys (raw packed input) is dead after the unique step: reuse it for the transposed keys.
    GP_LAUNCH(transpose_keys_kernel, launch_blocks(cap, 256), 256, 0, stream, c->ukeys, c->meta, nb, cap, c->keys);
    u64 *tsorted = c->keys;
    // distinct keys are (src,dst)-sorted, so a stable sort on the dst half alone yields (dst,src) order
    GP_TRY(gp_radix_sort(&c->sort_ws, c->keys, nullptr, (const u32 *)&c->meta[GP_META_NUM_EDGES], cap, nb,
                         2 * nb, stream, &tsorted, nullptr));
    GP_LAUNCH(set_meta_kernel, 1, 1, 0, stream, c->meta, GP_META_IS_SYMMETRIC, 1);
    GP_LAUNCH(symmetric_check_kernel, launch_blocks(cap, 256), 256, 0, stream, c->ukeys, tsorted, cap, c->meta);
    const int64_t work = cap > n + 1 ? cap : n + 1;
    GP_LAUNCH(unpack_csr_kernel, launch_blocks(work, 256), 256, 0, stream, tsorted, c->meta, n, nb, cap, c->rowptr_in,
                                                                    c->col_in);
    GP_LAUNCH(set_meta_kernel, 1, 1, 0, stream, c->meta, GP_META_IN_BUILT, 1);
    GP_CUDA_CHECK(cudaGetLastError());
    c->in_built = true;
    return GP_OK;
}

extern "C" int gp_csr_info(gp_csr_t *c, gp_csr_info_t *info, gp_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(c != nullptr && info != nullptr, GP_ERR_INVALID, "gp_csr_info: NULL argument");
    GP_REQUIRE(c->built, GP_ERR_INVALID, "gp_csr_info: gp_csr_build has not been called");
    GP_TRY(gp_csr_ensure_in(c, stream));
    int meta[GP_META_WORDS];
    GP_CUDA_CHECK(cudaMemcpyAsync(meta, c->meta, sizeof(meta), cudaMemcpyDeviceToHost, stream));
    GP_CUDA_CHECK(cudaStreamSynchronize(stream));
    info->num_nodes = c->num_nodes;
    info->num_input_edges = c->num_input_edges;
    info->num_edges = (u32)meta[GP_META_NUM_EDGES];
    info->max_out_degree = meta[GP_META_MAX_DEGREE];
    info->is_symmetric = meta[GP_META_IS_SYMMETRIC];
    info->reserved = 0;
    GP_REQUIRE(!(meta[GP_META_ERROR] & GP_DEV_ERR_EDGE_RANGE), GP_ERR_INDEX_RANGE,
               "edge_index holds an entry outside [0, %lld)", (long long)c->num_nodes);
    return GP_OK;
}

extern "C" int gp_csr_export(gp_csr_t *c, int which, int32_t *d_rowptr, int32_t *d_col, gp_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(c != nullptr && d_rowptr != nullptr && d_col != nullptr, GP_ERR_INVALID,
               "gp_csr_export: NULL argument");
    GP_REQUIRE(c->built, GP_ERR_INVALID, "gp_csr_export: gp_csr_build has not been called");
    GP_REQUIRE(which == 0 || which == 1, GP_ERR_INVALID, "gp_csr_export: which must be 0 (out) or 1 (in)");
    if (which == 1) GP_TRY(gp_csr_ensure_in(c, stream));
    int m = 0;
    GP_CUDA_CHECK(cudaMemcpyAsync(&m, &c->meta[GP_META_NUM_EDGES], sizeof(int), cudaMemcpyDeviceToHost, stream));
    GP_CUDA_CHECK(cudaStreamSynchronize(stream));
    GP_CUDA_CHECK(cudaMemcpyAsync(d_rowptr, which ? c->rowptr_in : c->rowptr_out,
                                  (size_t)(c->num_nodes + 1) * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    if (m > 0)
        GP_CUDA_CHECK(cudaMemcpyAsync(d_col, which ? c->col_in : c->col_out, (size_t)(u32)m * sizeof(int),
                                      cudaMemcpyDeviceToDevice, stream));
    GP_CUDA_CHECK(cudaStreamSynchronize(stream));
    return GP_OK;
}
