// gp_sort.cu — onesweep LSD radix sort + fused unique-compaction (see gp_sort.cuh).
//
// Used by the stable top-k of the centrality samplers (utils.py:29-30,41-42).  (The CSR builder
// used it in its first version; it is now a counting sort by row + in-row sorts, gp_csr.cu.)
#include "gp_sort.cuh"

namespace {

constexpr u32 ST_FLAG_SHIFT = 30;
constexpr u32 ST_AGG = 1u << ST_FLAG_SHIFT;   // tile aggregate published
constexpr u32 ST_INC = 2u << ST_FLAG_SHIFT;   // inclusive prefix published
constexpr u32 ST_VAL = (1u << ST_FLAG_SHIFT) - 1u;

struct PassPlan {
    int npass;
    int shift[GP_SORT_MAX_PASSES];
    int width[GP_SORT_MAX_PASSES];
};

__device__ __forceinline__ u32 live_count(const u32 *d_n, u32 n_max)
{
    if (d_n == nullptr) return n_max;
    u32 n = *d_n;
    return n < n_max ? n : n_max;
}

// Exclusive scan of one u32 per thread over a 256-thread block.
__device__ __forceinline__ u32 block_excl_scan_256(u32 v, u32 *s_warp, u32 &total)
{
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
    u32 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 y = __shfl_up_sync(FULL_MASK, x, o);
        if (lane >= (u32)o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    u32 wbase = 0, tot = 0;
#pragma unroll
    for (u32 w = 0; w < GP_SORT_THREADS / 32; ++w) {
        u32 t = s_warp[w];
        if (w < warp) wbase += t;
        tot += t;
    }
    __syncthreads();
    total = tot;
    return wbase + x - v;
}

// One pass over the keys builds the digit histograms of every radix pass.
__global__ void __launch_bounds__(GP_SORT_THREADS)
radix_hist_kernel(const u64 *__restrict__ keys, const u32 *d_n, u32 n_max, PassPlan plan, u32 *hist)
{
    __shared__ u32 sh[GP_SORT_MAX_PASSES * 256];
    const u32 n = live_count(d_n, n_max);
    for (int j = threadIdx.x; j < plan.npass * 256; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u64 k = keys[i];
        for (int p = 0; p < plan.npass; ++p) {
            const u32 d = (u32)(k >> plan.shift[p]) & ((1u << plan.width[p]) - 1u);
            atomicAdd(&sh[p * 256 + d], 1u);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < plan.npass * 256; j += blockDim.x)
        if (sh[j]) atomicAdd(&hist[j], sh[j]);
}

// One radix pass: rank keys inside a tile (warp match-any multisplit), publish the
// tile's per-digit counts, resolve the global offsets by decoupled look-back over
// earlier tiles, then scatter through shared memory so global writes are runs.
template <bool HAS_VALS>
__global__ void __launch_bounds__(GP_SORT_THREADS)
onesweep_kernel(const u64 *__restrict__ kin, u64 *__restrict__ kout, const u32 *__restrict__ vin,
                u32 *__restrict__ vout, const u32 *d_n, u32 n_max, const u32 *__restrict__ ghist,
                u32 *status, u32 *tile_counter, int shift, int width)
{
    __shared__ u64 s_keys[GP_SORT_TILE];
    __shared__ u32 s_vals[HAS_VALS ? GP_SORT_TILE : 1];
    __shared__ u32 s_whist[(GP_SORT_THREADS / 32) * 256];
    __shared__ u32 s_lexcl[256];
    __shared__ u32 s_gbase[256];
    __shared__ u32 s_warp[GP_SORT_THREADS / 32];
    __shared__ u32 s_tile;

    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const u32 n = live_count(d_n, n_max);
    const u32 ntiles = (n + GP_SORT_TILE - 1) / GP_SORT_TILE;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);  // dynamic ids: earlier tiles are running
    for (int j = tid; j < (GP_SORT_THREADS / 32) * 256; j += GP_SORT_THREADS) s_whist[j] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    if (tile >= ntiles) return;

    const u32 mask = (1u << width) - 1u;
    const u32 base = tile * GP_SORT_TILE + warp * (32 * GP_SORT_IPT);
    u64 key[GP_SORT_IPT];
    u32 val[GP_SORT_IPT];
    u32 rank[GP_SORT_IPT];
#pragma unroll
    for (int r = 0; r < GP_SORT_IPT; ++r) {
        const u32 idx = base + r * 32 + lane;
        key[r] = idx < n ? kin[idx] : ~0ull;
        if (HAS_VALS) val[r] = idx < n ? vin[idx] : 0u;
    }
    u32 *wh = s_whist + warp * 256;
#pragma unroll
    for (int r = 0; r < GP_SORT_IPT; ++r) {
        const u32 idx = base + r * 32 + lane;
        const bool valid = idx < n;
        const u32 d = valid ? ((u32)(key[r] >> shift) & mask) : 0xFFFFFFFFu;
        const u32 peers = __match_any_sync(FULL_MASK, d);
        const u32 leader = __ffs(peers) - 1;
        u32 old = 0;
        if (lane == leader && valid) {
            old = wh[d];
            wh[d] = old + __popc(peers);
        }
        old = __shfl_sync(FULL_MASK, old, leader);
        rank[r] = old + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();

    // thread `tid` owns digit `tid`: exclusive prefix over the warps of this tile
    u32 count = 0;
#pragma unroll
    for (int w = 0; w < GP_SORT_THREADS / 32; ++w) {
        const u32 c = s_whist[w * 256 + tid];
        s_whist[w * 256 + tid] = count;
        count += c;
    }
    u32 excl_prev = 0;
    if (tid <= mask) {
        u32 *st = status + (size_t)tile * 256 + tid;
        if (tile == 0) {
            st_relaxed_u32(st, ST_INC | count);
        } else {
            st_relaxed_u32(st, ST_AGG | count);
            // decoupled look-back, 8 predecessors per round trip: all tiles of a small sort start
            // together, so a one-at-a-time walk would serialise ~sqrt(2*tiles) L2 latencies
            int p = (int)tile - 1;
            bool done = false;
            while (!done) {
                u32 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    v[i] = (p - i >= 0) ? ld_relaxed_u32(status + (size_t)(p - i) * 256 + tid) : ST_INC;
                int used = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (done || used != i) continue;
                    const u32 f = v[i] >> ST_FLAG_SHIFT;
                    if (f == 0) continue;  // not published yet: poll again from here
                    excl_prev += v[i] & ST_VAL;
                    used = i + 1;
                    if (f == 2u) done = true;
                }
                p -= used;
            }
            st_relaxed_u32(st, ST_INC | (excl_prev + count));
        }
    }
    u32 total;
    const u32 lexcl = block_excl_scan_256(count, s_warp, total);
    const u32 gexcl = block_excl_scan_256(ghist[tid], s_warp, total);
    s_lexcl[tid] = lexcl;
    s_gbase[tid] = gexcl + excl_prev - lexcl;
    __syncthreads();

#pragma unroll
    for (int r = 0; r < GP_SORT_IPT; ++r) {
        const u32 idx = base + r * 32 + lane;
        if (idx < n) {
            const u32 d = (u32)(key[r] >> shift) & mask;
            const u32 pos = s_lexcl[d] + wh[d] + rank[r];
            s_keys[pos] = key[r];
            if (HAS_VALS) s_vals[pos] = val[r];
        }
    }
    __syncthreads();
    const u32 tile_n = min((u32)GP_SORT_TILE, n - tile * GP_SORT_TILE);
    for (u32 i = tid; i < tile_n; i += GP_SORT_THREADS) {
        const u64 k = s_keys[i];
        const u32 d = (u32)(k >> shift) & mask;
        const u32 dst = s_gbase[d] + i;
        kout[dst] = k;
        if (HAS_VALS) vout[dst] = s_vals[i];
    }
}

// Fused "flag heads / scan / scatter" over a sorted array, single pass with look-back.
// status[0] = tile ticket counter, status[1 + t] = tile t's prefix word.
__global__ void __launch_bounds__(GP_SORT_THREADS)
unique_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, const u32 *d_n, u32 n_max,
              u32 *d_m, u32 *status)
{
    __shared__ u32 s_warp[GP_SORT_THREADS / 32];
    __shared__ u32 s_tile;
    __shared__ u32 s_prefix;
    const u32 tid = threadIdx.x;
    const u32 n = live_count(d_n, n_max);
    const u32 ntiles = (n + GP_SORT_TILE - 1) / GP_SORT_TILE;
    if (tid == 0) s_tile = atomicAdd(status, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    if (tile >= ntiles) {
        if (ntiles == 0 && tile == 0 && tid == 0) *d_m = 0;
        return;
    }
    const u32 first = tile * GP_SORT_TILE + tid * GP_SORT_IPT;
    u64 k[GP_SORT_IPT];
    u64 prev = 0;
    if (first > 0 && first < n) prev = in[first - 1];
    u32 cnt = 0, flags = 0;
#pragma unroll
    for (int i = 0; i < GP_SORT_IPT; ++i) {
        const u32 idx = first + i;
        k[i] = idx < n ? in[idx] : 0ull;
        const bool head = idx < n && (idx == 0 || k[i] != prev);
        prev = k[i];
        flags |= (head ? 1u : 0u) << i;
        cnt += head ? 1u : 0u;
    }
    u32 total;
    const u32 lexcl = block_excl_scan_256(cnt, s_warp, total);
    if (tid < 32) {
        // warp-wide look-back over 32 predecessors per round trip
        const u32 lane = tid;
        u32 *st = status + 1 + tile;
        u32 excl = 0;
        bool done = tile == 0;
        if (lane == 0) st_relaxed_u32(st, (tile == 0 ? ST_INC : ST_AGG) | total);
        int p = (int)tile - 1;
        while (!done) {
            const int idx = p - (int)lane;
            const u32 v = idx >= 0 ? ld_relaxed_u32(status + 1 + idx) : ST_INC;
            const u32 f = v >> ST_FLAG_SHIFT;
            const u32 empty = __ballot_sync(FULL_MASK, f == 0);
            const u32 inc = __ballot_sync(FULL_MASK, f == 2u);
            const int first_empty = empty ? __ffs(empty) - 1 : 32;
            const int first_inc = inc ? __ffs(inc) - 1 : 32;
            const int take = first_inc < first_empty ? first_inc + 1 : first_empty;  // usable prefix of the window
            excl += __reduce_add_sync(FULL_MASK, (int)lane < take ? (v & ST_VAL) : 0u);
            p -= take;
            if (first_inc < first_empty) done = true;
        }
        if (lane == 0) {
            if (tile != 0) st_relaxed_u32(st, ST_INC | (excl + total));
            s_prefix = excl;
            if (tile == ntiles - 1) *d_m = excl + total;
        }
    }
    __syncthreads();
    u32 pos = s_prefix + lexcl;
#pragma unroll
    for (int i = 0; i < GP_SORT_IPT; ++i)
        if (flags & (1u << i)) out[pos++] = k[i];
}

}  // namespace

int gp_sort_workspace_create(GpSortWorkspace *ws, int64_t capacity, bool with_values)
{
    GP_REQUIRE(capacity >= 0 && capacity < (int64_t)ST_VAL, GP_ERR_UNSUPPORTED,
               "sort capacity %lld exceeds the 2^30-1 element limit", (long long)capacity);
    ws->capacity = capacity;
    ws->with_values = with_values;
    ws->max_tiles = gp_ceil_div(capacity > 0 ? capacity : 1, GP_SORT_TILE);
    const size_t cap = (size_t)(capacity > 0 ? capacity : 1);
    GP_CUDA_CHECK(cudaMalloc(&ws->keys_alt, cap * sizeof(u64)));
    if (with_values) GP_CUDA_CHECK(cudaMalloc(&ws->vals_alt, cap * sizeof(u32)));
    ws->scratch_bytes = sizeof(u32) * ((size_t)GP_SORT_MAX_PASSES * 256 + GP_SORT_MAX_PASSES +
                                       (size_t)GP_SORT_MAX_PASSES * ws->max_tiles * 256);
    GP_CUDA_CHECK(cudaMalloc(&ws->scratch, ws->scratch_bytes));
    return GP_OK;
}

void gp_sort_workspace_free(GpSortWorkspace *ws)
{
    if (ws->keys_alt) cudaFree(ws->keys_alt);
    if (ws->vals_alt) cudaFree(ws->vals_alt);
    if (ws->scratch) cudaFree(ws->scratch);
    *ws = GpSortWorkspace();
}

int gp_radix_sort(GpSortWorkspace *ws, u64 *keys, u32 *vals, const u32 *d_n, int64_t n_max,
                  int bit_lo, int bit_hi, cudaStream_t stream, u64 **keys_out, u32 **vals_out)
{
    *keys_out = keys;
    if (vals_out) *vals_out = vals;
    const int bits = bit_hi - bit_lo;
    if (bits <= 0 || n_max <= 0) return GP_OK;
    GP_REQUIRE(n_max <= ws->capacity, GP_ERR_INVALID, "radix sort: n_max %lld > workspace capacity %lld",
               (long long)n_max, (long long)ws->capacity);
    GP_REQUIRE(bit_lo >= 0 && bit_hi <= 64, GP_ERR_INVALID, "radix sort: bad bit range");
    GP_REQUIRE(vals == nullptr || ws->with_values, GP_ERR_INVALID, "radix sort: workspace has no value buffer");
    PassPlan plan;
    plan.npass = (bits + 7) / 8;
    {
        int pos = bit_lo, left = bits;
        for (int p = 0; p < plan.npass; ++p) {
            const int w = (left + (plan.npass - p) - 1) / (plan.npass - p);
            plan.shift[p] = pos;
            plan.width[p] = w;
            pos += w;
            left -= w;
        }
    }
    const int64_t tiles = gp_ceil_div(n_max, GP_SORT_TILE);
    u32 *hist = ws->scratch;
    u32 *counters = hist + GP_SORT_MAX_PASSES * 256;
    u32 *status = counters + GP_SORT_MAX_PASSES;
    const size_t used = sizeof(u32) * ((size_t)GP_SORT_MAX_PASSES * 256 + GP_SORT_MAX_PASSES +
                                       (size_t)plan.npass * tiles * 256);
    GP_CUDA_CHECK(cudaMemsetAsync(ws->scratch, 0, used, stream));
    int hist_blocks = (int)(tiles < (int64_t)gp_sm_count() * 4 ? tiles : (int64_t)gp_sm_count() * 4);
    GP_LAUNCH(radix_hist_kernel, hist_blocks, GP_SORT_THREADS, 0, stream, keys, d_n, (u32)n_max, plan, hist);
    u64 *ksrc = keys, *kdst = ws->keys_alt;
    u32 *vsrc = vals, *vdst = ws->vals_alt;
    for (int p = 0; p < plan.npass; ++p) {
        if (vals)
            GP_LAUNCH(onesweep_kernel<true>, (unsigned)tiles, GP_SORT_THREADS, 0, stream, ksrc, kdst, vsrc, vdst, d_n, (u32)n_max, hist + p * 256, status + (size_t)p * tiles * 256,
                counters + p, plan.shift[p], plan.width[p]);
        else
            GP_LAUNCH(onesweep_kernel<false>, (unsigned)tiles, GP_SORT_THREADS, 0, stream, ksrc, kdst, nullptr, nullptr, d_n, (u32)n_max, hist + p * 256,
                status + (size_t)p * tiles * 256, counters + p, plan.shift[p], plan.width[p]);
        u64 *tk = ksrc; ksrc = kdst; kdst = tk;
        u32 *tv = vsrc; vsrc = vdst; vdst = tv;
    }
    GP_CUDA_CHECK(cudaGetLastError());
    *keys_out = ksrc;
    if (vals_out) *vals_out = vsrc;
    return GP_OK;
}

int gp_unique_sorted(const u64 *in, u64 *out, const u32 *d_n, int64_t n_max, u32 *d_m, u32 *status,
                     cudaStream_t stream)
{
    const int64_t tiles = gp_ceil_div(n_max > 0 ? n_max : 1, GP_SORT_TILE);
    GP_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(u32) * (size_t)(tiles + 1), stream));
    GP_LAUNCH(unique_kernel, (unsigned)tiles, GP_SORT_THREADS, 0, stream, in, out, d_n, (u32)(n_max > 0 ? n_max : 0),
                                                                   d_m, status);
    GP_CUDA_CHECK(cudaGetLastError());
    return GP_OK;
}
