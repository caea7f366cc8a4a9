// gp_csr.cu — device CSR builder: the to_networkx(data) replacement (reference utils.py:121,
// also :27,33,39,45,51,57).  DiGraph semantics: nodes 0..N-1, parallel edges collapse,
// self-loops stay, no symmetrisation unless GP_CSR_SYMMETRIZE.
//
// The build is a counting sort by source row followed by a small sort inside every row:
//   count   cnt[src]++ for every column of edge_index                       (L2 reductions)
//   scan    raw_ptr = exclusive prefix of cnt (single-pass chained scan)    cursor = raw_ptr
//   scatter raw_col[atomicAdd(&cursor[src], 1)] = dst                       (arbitrary order inside a row)
//   rowsort every row segment sorted ascending and de-duplicated in place; rows of <= 32 edges in
//           registers by one warp (a sorting network over shuffles), longer rows by one CTA in
//           shared memory (a node bitmap for long rows, a sorting network otherwise) or in place in
//           global memory; deg[r] = distinct count
//   scan    nine prefix sums over the rows in ONE chained scan: edges (-> E'), work-list entries of
//           each of the seven degree classes (-> position of the row in the MS-BFS work list), hub
//           rows (-> index of the row's accumulator)
//   desc    the row's work-list descriptor(s)
// The CSR stays "gapped": row r is col[row_start[r] .. row_start[r] + deg[r]) and the slack left by
// dropped duplicates is never squeezed out (gp_csr_export compacts on request), so no column is
// moved twice.  The result is deterministic (rows ascending and unique, work list in ascending node
// order per class) although the scatter is not.  Seven launches over E + N ints replace the seven
// radix passes over 8-byte keys of the first version.  Everything is stream-ordered and
// graph-capturable; counts stay on the device.
//
// The in-edge CSR (PageRank pull, in-degrees) is the same pipeline run on the transposed edges of
// the finished CSR, built on demand.
#include "gp_internal.h"

#include <algorithm>
#include <new>

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;  // elements per tile
constexpr int ROW_CH = 9;                           // channels of the row scan (see RowScanIo)
constexpr int BIG_THREADS = 256;
constexpr int BIG_SMEM_ELEMS = 8192;                // longest row sorted in shared memory

int launch_blocks(int64_t work, int threads)
{
    int64_t b = gp_ceil_div(work > 0 ? work : 1, threads);
    const int64_t cap = (int64_t)gp_sm_count() * 32;  // latency-bound (atomics, dependent loads): a thread per item
    return (int)(b < cap ? b : cap);
}

// Grid for warp-per-row kernels.  They are chains of dependent loads per row, so every row gets its
// own warp while that still fits a few waves of resident CTAs.
int row_blocks(int64_t rows)
{
    int64_t b = gp_ceil_div(rows > 0 ? rows : 1, 8);  // 8 warps per 256-thread CTA
    const int64_t cap = (int64_t)gp_sm_count() * 64;
    return (int)(b < cap ? b : cap);
}

// ---------------------------------------------------------------- count / scatter
// Four consecutive edges per thread, read as two 16-byte loads from each row of edge_index: one resident wave covers a
// million edges, and every thread has its eight atomics in flight together (a thread per edge ran four waves, each a
// dependent load -> atomic round trip).  `vec` = both rows are 16-byte aligned (E even and the base aligned).
__device__ __forceinline__ void load_edges4(const long long *__restrict__ ei, long long e, long long i0, int vec,
                                            long long (&s)[4], long long (&d)[4], int &cnt)
{
    cnt = (int)(e - i0 < 4 ? e - i0 : 4);
    if (vec && cnt == 4) {
        const longlong2 a = __ldcs(reinterpret_cast<const longlong2 *>(ei + i0));
        const longlong2 b = __ldcs(reinterpret_cast<const longlong2 *>(ei + i0) + 1);
        const longlong2 c = __ldcs(reinterpret_cast<const longlong2 *>(ei + e + i0));
        const longlong2 f = __ldcs(reinterpret_cast<const longlong2 *>(ei + e + i0) + 1);
        s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y;
        d[0] = c.x; d[1] = c.y; d[2] = f.x; d[3] = f.y;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            s[q] = q < cnt ? ei[i0 + q] : 0;
            d[q] = q < cnt ? ei[e + i0 + q] : 0;
        }
    }
}

__global__ void __launch_bounds__(256)
count_edges_kernel(const long long *__restrict__ ei, long long e, long long n, int symmetrize, int vec,
                   int *__restrict__ cnt, int *meta)
{
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    bool bad = false;
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 < e; i0 += stride) {
        long long s[4], d[4];
        int c;
        load_edges4(ei, e, i0, vec, s, d, c);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q >= c) break;
            if (s[q] < 0 || d[q] < 0 || s[q] >= n || d[q] >= n) {
                bad = true;
                continue;
            }
            atomicAdd(cnt + s[q], 1);
            if (symmetrize) atomicAdd(cnt + d[q], 1);
        }
    }
    if (__any_sync(FULL_MASK, bad) && lane_id() == 0) atomicOr(&meta[GP_META_ERROR], GP_DEV_ERR_EDGE_RANGE);
}

__global__ void __launch_bounds__(256)
scatter_edges_kernel(const long long *__restrict__ ei, long long e, long long n, int symmetrize, int vec,
                     int *__restrict__ cursor, int *__restrict__ raw_col)
{
    gp_pdl_enter();
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 < e; i0 += stride) {
        long long s[4], d[4];
        int c, pos[4], pos2[4];
        load_edges4(ei, e, i0, vec, s, d, c);
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // the four returning atomics are in flight together
            ok[q] = q < c && !(s[q] < 0 || d[q] < 0 || s[q] >= n || d[q] >= n);  // bad entries: latched by count_edges_kernel
            pos[q] = ok[q] ? atomicAdd(cursor + s[q], 1) : 0;
            pos2[q] = (ok[q] && symmetrize) ? atomicAdd(cursor + d[q], 1) : 0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (ok[q]) {
                raw_col[pos[q]] = (int)d[q];
                if (symmetrize) raw_col[pos2[q]] = (int)s[q];
            }
        }
    }
}

// Transposed counting / scatter from a finished CSR (warp per row).
__global__ void count_in_kernel(const int *__restrict__ row_start, const int *__restrict__ deg,
                                const int *__restrict__ col, long long n, int *__restrict__ cnt)
{
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = lane_id();
    for (long long u = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n; u += warps) {
        const int s = row_start[u], t = s + deg[u];
        for (int j = s + lane; j < t; j += 32) atomicAdd(cnt + col[j], 1);
    }
}

__global__ void scatter_in_kernel(const int *__restrict__ row_start, const int *__restrict__ deg,
                                  const int *__restrict__ col, long long n, int *__restrict__ cursor,
                                  int *__restrict__ col_in)
{
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = lane_id();
    for (long long u = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n; u += warps) {
        const int s = row_start[u], t = s + deg[u];
        for (int j = s + lane; j < t; j += 32) col_in[atomicAdd(cursor + col[j], 1)] = (int)u;
    }
}

// ---------------------------------------------------------------- single-pass chained scan
// Status words of tile t: [t * STRIDE] flag (0 nothing, 1 aggregate, 2 inclusive prefix), then CHN
// aggregates, then CHN inclusive prefixes.  Tiles take their ids from a ticket counter, so every
// predecessor of a running tile is running or finished and the look-back cannot starve.
template <int CHN>
struct ScanStatus {
    static constexpr int STRIDE = 1 + 2 * CHN;
};

__device__ __forceinline__ int ld_acquire_i32(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_i32(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// IO: struct with
//   __device__ void load(long long i, int (&v)[CHN]) const      contribution of element i (v arrives zeroed)
//   __device__ void store(long long i, const int (&excl)[CHN], const int (&v)[CHN]) const
//   __device__ void finish(const int (&total)[CHN]) const        called once, by the last tile
template <int CHN, class IO, int IPT = SCAN_IPT>
__global__ void __launch_bounds__(SCAN_THREADS) chained_scan_kernel(IO io, long long count, int *status, int *ticket)
{
    constexpr int STRIDE = ScanStatus<CHN>::STRIDE;
    constexpr int SCAN_IPT = IPT;                         // items per thread of THIS instantiation
    constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;    // (shadow the file-level defaults)
    __shared__ int s_tile;
    __shared__ int s_warp[SCAN_THREADS / 32][CHN];
    __shared__ int s_prefix[CHN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    gp_pdl_enter();
    if (tid == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    const long long ntiles = (count + SCAN_TILE - 1) / SCAN_TILE;
    if (tile >= ntiles) return;

    const long long base = (long long)tile * SCAN_TILE + (long long)tid * SCAN_IPT;
    int v[SCAN_IPT][CHN], tsum[CHN];
#pragma unroll
    for (int c = 0; c < CHN; ++c) tsum[c] = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
#pragma unroll
        for (int c = 0; c < CHN; ++c) v[k][c] = 0;
        if (base + k < count) io.load(base + k, v[k]);
#pragma unroll
        for (int c = 0; c < CHN; ++c) tsum[c] += v[k][c];
    }
    // block-wide exclusive scan of the per-thread sums, channel by channel
    int texcl[CHN];
#pragma unroll
    for (int c = 0; c < CHN; ++c) {
        int x = tsum[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL_MASK, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp][c] = x;
        texcl[c] = x - tsum[c];
    }
    __syncthreads();
    int agg[CHN];
#pragma unroll
    for (int c = 0; c < CHN; ++c) {
        int wbase = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) {
            const int t = s_warp[w][c];
            if (w < warp) wbase += t;
            tot += t;
        }
        texcl[c] += wbase;
        agg[c] = tot;
    }
    // publish the aggregate, then warp 0 looks back over 32 predecessors per round trip
    int *st = status + (size_t)tile * STRIDE;
    if (warp == 0) {
        if (tile > 0) {
            if (lane < CHN) {
                int mine = 0;
#pragma unroll
                for (int c = 0; c < CHN; ++c)
                    if (c == lane) mine = agg[c];
                st[1 + lane] = mine;
                __threadfence();
            }
            __syncwarp();
            if (lane == 0) st_release_i32(st, 1);
        }
        int run[CHN];
#pragma unroll
        for (int c = 0; c < CHN; ++c) run[c] = 0;
        for (int t = tile - 1; t >= 0; t -= 32) {
            const int idx = t - lane;
            const int *pt = status + (size_t)(idx >= 0 ? idx : 0) * STRIDE;
            int f = idx >= 0 ? ld_acquire_i32(pt) : 2;  // tiles before the first count as a zero inclusive prefix
            while (__any_sync(FULL_MASK, f == 0))
                if (f == 0) f = ld_acquire_i32(pt);
            const u32 inc_mask = __ballot_sync(FULL_MASK, f == 2);
            const int first = inc_mask ? __ffs(inc_mask) - 1 : 32;  // nearest predecessor with an inclusive prefix
#pragma unroll
            for (int c = 0; c < CHN; ++c) {
                int val = 0;
                if (idx >= 0 && lane <= first) val = pt[1 + (lane == first ? CHN : 0) + c];
                run[c] += __reduce_add_sync(FULL_MASK, val);
            }
            if (inc_mask) break;
        }
        if (lane < CHN) {
            int mine = 0, r = 0;
#pragma unroll
            for (int c = 0; c < CHN; ++c)
                if (c == lane) {
                    mine = agg[c];
                    r = run[c];
                }
            st[1 + CHN + lane] = r + mine;
            s_prefix[lane] = r;
            __threadfence();
        }
        __syncwarp();
        if (lane == 0) st_release_i32(st, 2);
    }
    __syncthreads();
    int excl[CHN];
#pragma unroll
    for (int c = 0; c < CHN; ++c) excl[c] = s_prefix[c] + texcl[c];
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        if (base + k < count) io.store(base + k, excl, v[k]);
#pragma unroll
        for (int c = 0; c < CHN; ++c) excl[c] += v[k][c];
    }
    if (tile == ntiles - 1 && tid == 0) {
        int total[CHN];
#pragma unroll
        for (int c = 0; c < CHN; ++c) total[c] = s_prefix[c] + agg[c];
        io.finish(total);
    }
}

// cnt[0..n) -> ptr[0..n] (exclusive prefix, ptr[n] = total) and a copy in cursor[0..n).
struct PtrScanIo {
    const int *cnt;
    int *ptr;
    int *cursor;
    long long n;
    __device__ void load(long long i, int (&v)[1]) const
    {
        if (i < n) v[0] = cnt[i];
    }
    __device__ void store(long long i, const int (&excl)[1], const int (&)[1]) const
    {
        ptr[i] = excl[0];
        if (i < n) cursor[i] = excl[0];
    }
    __device__ void finish(const int (&)[1]) const {}
};

// The same for the first scan of gp_csr_build; it also queues the rows of more than 128 raw edges for the CTA-wide
// row sort, which can then run BESIDE the sort of the short rows instead of after it.
struct PtrScanBuildIo {
    const int *cnt;
    int *ptr;
    int *cursor;
    long long n;
    int *biglist;
    int *medlist;
    int *meta;
    __device__ void load(long long i, int (&v)[1]) const
    {
        if (i < n) v[0] = cnt[i];
    }
    __device__ void store(long long i, const int (&excl)[1], const int (&v)[1]) const
    {
        ptr[i] = excl[0];
        if (i < n) {
            cursor[i] = excl[0];
            if (v[0] > 128) biglist[atomicAdd(&meta[GP_META_NUM_BIG_ROWS], 1)] = (int)i;
            else if (v[0] > 16) medlist[atomicAdd(&meta[GP_META_NUM_MED_ROWS], 1)] = (int)i;
        }
    }
    __device__ void finish(const int (&)[1]) const {}
};

__device__ __forceinline__ int degree_class(int d)
{
    // 0: hub rows (cut into GP_CHUNK_EDGES chunks), then G = 32, 16, 8, 4, 2, 1 slots of GP_SLOT_EDGES edges
    return d > GP_CHUNK_EDGES ? 0 : d > 64 ? 1 : d > 32 ? 2 : d > 16 ? 3 : d > 8 ? 4 : d > 4 ? 5 : 6;
}

// Row scan: channel 0 edges, 1..7 work-list entries of class 0..6, 8 hub rows.  The scan writes the work-list
// descriptors itself: the descriptors of a class live in a region whose START is fixed when the handle is created
// (desc_off), so a row's position is known from its exclusive prefix alone and no second pass over the rows is needed.
//   desc = {row, first edge, count | chunks << 8 | first << 30, hub index or -1}, one per row, hub rows one per
//   GP_CHUNK_EDGES-edge chunk.
struct RowScanIo {
    const int *deg;
    const int *row_start;
    int4 *desc;
    int off[GP_NUM_CLASSES];
    int *meta;
    long long n;
    __device__ void load(long long i, int (&v)[ROW_CH]) const
    {
        const int d = deg[i];
        const int c = degree_class(d);
        v[0] = d;
#pragma unroll
        for (int k = 0; k < GP_NUM_CLASSES; ++k)
            if (k == c) v[1 + k] = c == 0 ? (d + GP_CHUNK_EDGES - 1) / GP_CHUNK_EDGES : 1;
        v[8] = c == 0 ? 1 : 0;
    }
    __device__ void store(long long i, const int (&excl)[ROW_CH], const int (&v)[ROW_CH]) const
    {
        const int d = v[0];
        const int c = degree_class(d);
        int r = 0, base = 0;
#pragma unroll
        for (int k = 0; k < GP_NUM_CLASSES; ++k)
            if (k == c) {
                r = excl[1 + k];
                base = off[k];
            }
        const int s = row_start[i];
        if (c != 0) {
            desc[base + r] = make_int4((int)i, s, d, -1);
        } else {
            const int nch = (d + GP_CHUNK_EDGES - 1) / GP_CHUNK_EDGES;
            for (int k = 0; k < nch; ++k) {
                const int cnt = min(GP_CHUNK_EDGES, d - k * GP_CHUNK_EDGES);
                desc[base + r + k] = make_int4((int)i, s + k * GP_CHUNK_EDGES, cnt | (nch << 8) | (k == 0 ? 1 << 30 : 0), excl[8]);
            }
        }
    }
    __device__ void finish(const int (&total)[ROW_CH]) const
    {
        meta[GP_META_NUM_EDGES] = total[0];
        meta[GP_META_NUM_HUB_ROWS] = total[8];
        // class c: descriptors [ent_base[c], ent_base[c+1]), G(c) slots each, regions tile aligned
        int ent = 0, slot = 0;
        for (int c = 0; c < GP_NUM_CLASSES; ++c) {
            const int g = c <= 1 ? 32 : (32 >> (c - 1));
            meta[GP_META_ENT_BASE + c] = ent;
            meta[GP_META_SLOT_BASE + c] = slot;
            ent += total[1 + c];
            slot += (int)(((long long)total[1 + c] * g + GP_SLOT_ALIGN - 1) / GP_SLOT_ALIGN * GP_SLOT_ALIGN);
        }
        meta[GP_META_ENT_BASE + GP_NUM_CLASSES] = ent;
        meta[GP_META_SLOT_BASE + GP_NUM_CLASSES] = slot;
    }
};

// ---------------------------------------------------------------- sort inside the rows
// Sorting network with every comparator pointing the same way (the lower index keeps the smaller
// key): for k = 2, 4, ...: partner = i ^ (k - 1) ("flip"), then partner = i ^ j for j = k/4, ..., 1
// ("disperse").  Indices >= len act as +infinity and are never moved, so any length sorts in place.
__device__ __forceinline__ int warp_sort32(int v, int lane)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        {
            const int o = __shfl_xor_sync(FULL_MASK, v, k - 1);
            v = (lane & (k >> 1)) ? max(v, o) : min(v, o);
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            const int o = __shfl_xor_sync(FULL_MASK, v, j);
            v = (lane & j) ? max(v, o) : min(v, o);
        }
    }
    return v;
}

// The same network over E * 32 keys held E per lane: key i lives in register i / 32 of lane i % 32,
// so partners at distance < 32 are one shuffle away and partners at distance >= 32 sit in another
// register (of the mirrored lane for a flip step).
template <int E>
__device__ __forceinline__ void warp_sort_multi(int (&v)[E], int lane)
{
#pragma unroll
    for (int k = 2; k <= 32 * E; k <<= 1) {
        // flip: partner = i ^ (k - 1)
        if (k <= 32) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int o = __shfl_xor_sync(FULL_MASK, v[e], k - 1);
                v[e] = (lane & (k >> 1)) ? max(v[e], o) : min(v[e], o);
            }
        } else {
            const int em = (k >> 5) - 1;  // register index bits flipped by this block size
            int o[E];
#pragma unroll
            for (int e = 0; e < E; ++e) o[e] = __shfl_xor_sync(FULL_MASK, v[e ^ em], 31);
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = (e & (k >> 6)) ? max(v[e], o[e]) : min(v[e], o[e]);
        }
        // disperse: partner = i ^ j
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j >= 32) {
                const int ej = j >> 5;
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (!(e & ej)) {
                        const int lo = min(v[e], v[e | ej]), hi = max(v[e], v[e | ej]);
                        v[e] = lo;
                        v[e | ej] = hi;
                    }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int o = __shfl_xor_sync(FULL_MASK, v[e], j);
                    v[e] = (lane & j) ? max(v[e], o) : min(v[e], o);
                }
            }
        }
    }
}

// Sort + unique of one row of <= 32 * E edges by one warp, in registers; returns the distinct count.
template <int E>
__device__ __forceinline__ int warp_row_sort_unique(int *colbuf, int s, int len, int lane)
{
    int v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = e * 32 + lane < len ? colbuf[s + e * 32 + lane] : 0x7FFFFFFF;
    if constexpr (E == 1) v[0] = warp_sort32(v[0], lane);
    else warp_sort_multi<E>(v, lane);
    int base = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        int prev = __shfl_up_sync(FULL_MASK, v[e], 1);
        if (e > 0) {
            const int last = __shfl_sync(FULL_MASK, v[e - 1], 31);
            if (lane == 0) prev = last;
        }
        const bool head = e * 32 + lane < len && ((e == 0 && lane == 0) || v[e] != prev);
        const u32 heads = __ballot_sync(FULL_MASK, head);
        // in place is safe: every key of the row is already in registers
        if (head) colbuf[s + base + __popc(heads & ((1u << lane) - 1u))] = v[e];
        base += __popc(heads);
    }
    return base;
}

// Sort + unique of one row of <= 8 edges by ONE thread (19-comparator network in registers).  Most rows
// of the named graphs are this short, and a warp of such rows reads one contiguous stretch of colbuf.
__device__ __forceinline__ int thread_row_sort_unique(int *colbuf, int s, int len)
{
    int v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < len ? colbuf[s + i] : 0x7FFFFFFF;
#define GP_CE(a, b)                              \
    {                                            \
        const int lo = min(v[a], v[b]);          \
        v[b] = max(v[a], v[b]);                  \
        v[a] = lo;                               \
    }
    GP_CE(0, 1) GP_CE(2, 3) GP_CE(4, 5) GP_CE(6, 7)
    GP_CE(0, 2) GP_CE(1, 3) GP_CE(4, 6) GP_CE(5, 7)
    GP_CE(1, 2) GP_CE(5, 6) GP_CE(0, 4) GP_CE(3, 7)
    GP_CE(1, 5) GP_CE(2, 6)
    GP_CE(1, 4) GP_CE(3, 6)
    GP_CE(2, 4) GP_CE(3, 5)
    GP_CE(3, 4)
#undef GP_CE
    int d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < len && (i == 0 || v[i] != v[i - 1])) {
            colbuf[s + d] = v[i];
            ++d;
        }
    }
    return d;
}

// The same for rows of 9..16 edges (63-comparator odd-even merge network).  On the Flickr-shaped graph a third of
// the rows are longer than 8 edges; walking them one at a time per warp was the longest part of the build.
__device__ __forceinline__ int thread_row_sort_unique16(int *colbuf, int s, int len)
{
    int v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = i < len ? colbuf[s + i] : 0x7FFFFFFF;
#define GP_CE(a, b)                              \
    {                                            \
        const int lo = min(v[a], v[b]);          \
        v[b] = max(v[a], v[b]);                  \
        v[a] = lo;                               \
    }
    GP_CE(0, 1) GP_CE(2, 3) GP_CE(0, 2) GP_CE(1, 3) GP_CE(1, 2) GP_CE(4, 5) GP_CE(6, 7) GP_CE(4, 6)
    GP_CE(5, 7) GP_CE(5, 6) GP_CE(0, 4) GP_CE(2, 6) GP_CE(2, 4) GP_CE(1, 5) GP_CE(3, 7) GP_CE(3, 5)
    GP_CE(1, 2) GP_CE(3, 4) GP_CE(5, 6) GP_CE(8, 9) GP_CE(10, 11) GP_CE(8, 10) GP_CE(9, 11)
    GP_CE(9, 10) GP_CE(12, 13) GP_CE(14, 15) GP_CE(12, 14) GP_CE(13, 15) GP_CE(13, 14) GP_CE(8, 12)
    GP_CE(10, 14) GP_CE(10, 12) GP_CE(9, 13) GP_CE(11, 15) GP_CE(11, 13) GP_CE(9, 10) GP_CE(11, 12)
    GP_CE(13, 14) GP_CE(0, 8) GP_CE(4, 12) GP_CE(4, 8) GP_CE(2, 10) GP_CE(6, 14) GP_CE(6, 10)
    GP_CE(2, 4) GP_CE(6, 8) GP_CE(10, 12) GP_CE(1, 9) GP_CE(5, 13) GP_CE(5, 9) GP_CE(3, 11)
    GP_CE(7, 15) GP_CE(7, 11) GP_CE(3, 5) GP_CE(7, 9) GP_CE(11, 13) GP_CE(1, 2) GP_CE(3, 4) GP_CE(5, 6)
    GP_CE(7, 8) GP_CE(9, 10) GP_CE(11, 12) GP_CE(13, 14)
#undef GP_CE
    int d = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (i < len && (i == 0 || v[i] != v[i - 1])) {
            colbuf[s + d] = v[i];
            ++d;
        }
    }
    return d;
}

// Rows of <= 128 edges.  Pass 1: a thread per row for rows of <= 16 edges.  Pass 2: every warp walks the
// longer rows among its own 32 (registers only: 1, 2 or 4 keys per lane).  Rows above 128 edges are
// queued for rowsort_big_kernel.
__global__ void __launch_bounds__(256)
rowsort_small_kernel(const int *__restrict__ ptr, const int *len_in, int *__restrict__ colbuf, long long n, int *deg,
                     int *__restrict__ biglist, int *meta, int max_word, int collect_big)
{
    const int lane = lane_id();
    const long long nblk = (n + 255) / 256;
    int mx = 0;
    gp_pdl_enter();
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long r = blk * 256 + threadIdx.x;
        int s = 0, len = 0;
        if (r < n) {
            s = ptr[r];
            len = len_in[r];
        }
        if (len > 0 && len <= 8) {
            const int d = thread_row_sort_unique(colbuf, s, len);
            if (d != len) deg[r] = d;
            mx = max(mx, d);
        } else if (len > 8 && len <= 16) {
            const int d = thread_row_sort_unique16(colbuf, s, len);
            if (d != len) deg[r] = d;
            mx = max(mx, d);
        }
        u32 longer = collect_big ? __ballot_sync(FULL_MASK, len > 16) : 0u;  // build mode: queued rows, other kernel
        while (longer) {
            const int src = __ffs(longer) - 1;
            longer &= longer - 1;
            const int rs = __shfl_sync(FULL_MASK, s, src), rlen = __shfl_sync(FULL_MASK, len, src);
            const long long rr = r - lane + src;
            if (rlen > 128) {  // queued for rowsort_big_kernel (by the first scan of gp_csr_build when !collect_big)
                if (collect_big && lane == 0) biglist[atomicAdd(&meta[GP_META_NUM_BIG_ROWS], 1)] = (int)rr;
                continue;
            }
            int d;
            if (rlen <= 32) d = warp_row_sort_unique<1>(colbuf, rs, rlen, lane);
            else if (rlen <= 64) d = warp_row_sort_unique<2>(colbuf, rs, rlen, lane);
            else d = warp_row_sort_unique<4>(colbuf, rs, rlen, lane);
            if (lane == 0 && d != rlen) deg[rr] = d;
            mx = max(mx, d);
        }
    }
    // one same-address atomic per warp would serialise in L2: only warps that raise the maximum issue one
    mx = __reduce_max_sync(FULL_MASK, mx);
    if (lane == 0 && mx > 0 && mx > *(volatile int *)&meta[max_word]) atomicMax(&meta[max_word], mx);
}

// One step of the network over a[0..len): `sh` = log2 of the half block (flip) or of the stride
// (disperse).  Thread t handles comparator pairs t, t + T, ...; four pairs are loaded before any is
// written so the loads overlap.
template <bool FLIP>
__device__ __forceinline__ void network_step(int *a, int len, int pairs, int sh, int tid, int nthreads)
{
    const int h = 1 << sh;
    for (int p0 = tid; p0 < pairs; p0 += 4 * nthreads) {
        int i[4], o[4], x[4], y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int p = p0 + q * nthreads;
            const int blk = p >> sh, off = p & (h - 1);
            i[q] = (blk << (sh + 1)) + off;
            o[q] = FLIP ? (blk << (sh + 1)) + 2 * h - 1 - off : i[q] + h;
            if (p >= pairs) o[q] = len;  // inactive
            if (o[q] < len) {
                x[q] = a[i[q]];
                y[q] = a[o[q]];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (o[q] < len && x[q] > y[q]) {
                a[i[q]] = y[q];
                a[o[q]] = x[q];
            }
    }
}

// One CTA per queued row.  Long rows of a graph whose node bitmap fits in shared memory are sorted
// and de-duplicated by setting one bit per neighbour and reading the bits back in order; the others
// go through the sorting network (shared memory up to BIG_SMEM_ELEMS, else in place) and a head-flag
// scan.  Dynamic shared memory: max(BIG_SMEM_ELEMS ints, bitmap_words words).
__global__ void __launch_bounds__(BIG_THREADS)
rowsort_big_kernel(const int *__restrict__ ptr, const int *len_in, int *colbuf, int *deg,
                   const int *__restrict__ biglist, int *meta, int max_word, int bitmap_words,
                   const int *__restrict__ medlist)
{
    extern __shared__ __align__(16) int s_buf[];
    __shared__ int s_warp[BIG_THREADS / 32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    gp_pdl_enter();
    const int nbig = meta[GP_META_NUM_BIG_ROWS];
    for (int q = blockIdx.x; q < nbig; q += gridDim.x) {
        const int r = biglist[q];
        const int s = ptr[r], len = len_in[r];
        __syncthreads();  // previous row done with s_buf / s_carry
        if (bitmap_words > 0 && (long long)len * 32 >= bitmap_words) {
            // ---- bitmap path
            u32 *bm = reinterpret_cast<u32 *>(s_buf);
            for (int i = tid; i < bitmap_words; i += BIG_THREADS) bm[i] = 0;
            __syncthreads();
            for (int i = tid; i < len; i += BIG_THREADS) {
                const int v = colbuf[s + i];
                atomicOr(&bm[v >> 5], 1u << (v & 31));
            }
            __syncthreads();
            // thread t owns the contiguous words [t * wpt, (t + 1) * wpt)
            const int wpt = (bitmap_words + BIG_THREADS - 1) / BIG_THREADS;
            const int w0 = tid * wpt, w1 = min(bitmap_words, w0 + wpt);
            int cnt = 0;
            for (int w = w0; w < w1; ++w) cnt += __popc(bm[w]);
            int x = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL_MASK, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_warp[warp] = x;
            __syncthreads();
            int wbase = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < BIG_THREADS / 32; ++w) {
                const int t = s_warp[w];
                if (w < warp) wbase += t;
                tot += t;
            }
            int pos = s + wbase + x - cnt;
            for (int w = w0; w < w1; ++w) {
                u32 bits = bm[w];
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    colbuf[pos++] = (w << 5) + b;
                }
            }
            if (tid == 0) {
                if (tot != len) deg[r] = tot;
                if (tot > *(volatile int *)&meta[max_word]) atomicMax(&meta[max_word], tot);
            }
            continue;
        }
        // ---- network path
        const bool in_smem = len <= BIG_SMEM_ELEMS;
        int *a = in_smem ? s_buf : colbuf + s;
        if (in_smem)
            for (int i = tid; i < len; i += BIG_THREADS) s_buf[i] = colbuf[s + i];
        __syncthreads();
        int lg = 1;
        while ((1 << lg) < len) ++lg;  // network over 2^lg >= len virtual elements
        const int pairs = 1 << (lg - 1);
        for (int kb = 1; kb <= lg; ++kb) {  // block size k = 2^kb
            network_step<true>(a, len, pairs, kb - 1, tid, BIG_THREADS);
            __syncthreads();
            for (int jb = kb - 2; jb >= 0; --jb) {
                network_step<false>(a, len, pairs, jb, tid, BIG_THREADS);
                __syncthreads();
            }
        }
        // unique: out position = number of heads before i; chunks of BIG_THREADS with a running carry.
        // Writing in place is safe: position <= i, and a chunk is written only after it was read.
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < len; base += BIG_THREADS) {
            const int i = base + tid;
            int v = 0;
            bool head = false;
            if (i < len) {
                v = a[i];
                head = i == 0 || a[i - 1] != v;
            }
            const u32 heads = __ballot_sync(FULL_MASK, head);
            if (lane == 0) s_warp[warp] = __popc(heads);
            __syncthreads();  // every a[i], a[i-1] of this chunk has been read
            int wbase = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < BIG_THREADS / 32; ++w) {
                const int t = s_warp[w];
                if (w < warp) wbase += t;
                tot += t;
            }
            const int carry = s_carry;
            if (head) colbuf[s + carry + wbase + __popc(heads & ((1u << lane) - 1u))] = v;
            __syncthreads();
            if (tid == 0) s_carry = carry + tot;
            __syncthreads();
        }
        if (tid == 0) {
            if (s_carry != len) deg[r] = s_carry;
            if (s_carry > *(volatile int *)&meta[max_word]) atomicMax(&meta[max_word], s_carry);
        }
    }
    // ---- rows of 17..128 raw edges (gp_csr_build only): one warp each, in registers.  The CTAs that had no long row
    // take them (all CTAs when every CTA had one), so they run beside the long rows and beside rowsort_small_kernel.
    if (medlist != nullptr) {
        const int nmed = meta[GP_META_NUM_MED_ROWS];
        const int busy = nbig < (int)gridDim.x ? nbig : 0;  // CTAs [0, busy) are still sorting a long row
        const int wpc = BIG_THREADS / 32;
        const int nw = ((int)gridDim.x - busy) * wpc;
        int mx = 0;
        if ((int)blockIdx.x >= busy) {
            for (int q = ((int)blockIdx.x - busy) * wpc + warp; q < nmed; q += nw) {
                const int r = medlist[q];
                const int rs = ptr[r], rlen = len_in[r];
                int d;
                if (rlen <= 32) d = warp_row_sort_unique<1>(colbuf, rs, rlen, lane);
                else if (rlen <= 64) d = warp_row_sort_unique<2>(colbuf, rs, rlen, lane);
                else d = warp_row_sort_unique<4>(colbuf, rs, rlen, lane);
                if (lane == 0 && d != rlen) deg[r] = d;
                mx = max(mx, d);
            }
        }
        if (lane == 0 && mx > 0 && mx > *(volatile int *)&meta[max_word]) atomicMax(&meta[max_word], mx);
    }
}

// is_symmetric: every row of the out-edge CSR equals the same row of the in-edge CSR (warp per row).
__global__ void compare_csr_kernel(const int *__restrict__ row_start, const int *__restrict__ deg,
                                   const int *__restrict__ col, const int *__restrict__ rowptr_in,
                                   const int *__restrict__ col_in, long long n, int *meta)
{
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = lane_id();
    bool diff = false;
    for (long long u = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n; u += warps) {
        const int s = row_start[u], d = deg[u], si = rowptr_in[u];
        if (rowptr_in[u + 1] - si != d) {
            diff = true;
            continue;
        }
        for (int j = lane; j < d; j += 32)
            if (col[s + j] != col_in[si + j]) diff = true;
    }
    if (__any_sync(FULL_MASK, diff) && lane == 0) atomicExch(&meta[GP_META_IS_SYMMETRIC], 0);
}

// Compacting copy for gp_csr_export (warp per row).
__global__ void gather_rows_kernel(const int *__restrict__ row_start, const int *__restrict__ deg,
                                   const int *__restrict__ col, const int *__restrict__ rowptr, long long n,
                                   int *__restrict__ out)
{
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = lane_id();
    for (long long u = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n; u += warps) {
        const int s = row_start[u], d = deg[u], t = rowptr[u];
        for (int j = lane; j < d; j += 32) out[t + j] = col[s + j];
    }
}

__global__ void set_meta_kernel(int *meta, int idx, int value) { meta[idx] = value; }

int scan_tiles(int64_t count) { return (int)gp_ceil_div(count > 0 ? count : 1, SCAN_TILE); }

// Sorts (and de-duplicates) the row segments [ptr[r], ptr[r] + len[r]) of colbuf in place; len[r] is
// overwritten with the distinct count where it shrinks; meta[max_word] = max distinct count.
int sort_rows(gp_csr *c, const int *ptr, int *len, int *colbuf, int max_word, bool reset_queue, cudaStream_t stream)
{
    const int64_t n = c->num_nodes;
    if (reset_queue) GP_LAUNCH(set_meta_kernel, 1, 1, 0, stream, c->meta, GP_META_NUM_BIG_ROWS, 0);
    // programmatic launches: inside gp_csr_build they follow kernels of the same chain (after a memset or a
    // one-thread kernel the attribute simply has no effect)
    GP_LAUNCH_PDL(rowsort_small_kernel, launch_blocks(n, 256), 256, 0, stream, ptr, (const int *)len, colbuf, (long long)n, len,
                  c->biglist, c->meta, max_word, 1);
    GP_LAUNCH_PDL(rowsort_big_kernel, gp_sm_count() * 4, BIG_THREADS, (size_t)c->big_smem_bytes, stream, ptr,
                  (const int *)len, colbuf, len, (const int *)c->biglist, c->meta, max_word, c->bitmap_words,
                  (const int *)nullptr);
    return GP_OK;
}

// The same with the long rows already queued (PtrScanBuildIo): the two kernels touch disjoint rows, so the CTA-wide
// sort of the long rows runs on a side stream (a parallel branch of a captured graph) beside the sort of the short ones.
int sort_rows_forked(gp_csr *c, const int *ptr, int *len, int *colbuf, int max_word, cudaStream_t stream)
{
    const int64_t n = c->num_nodes;
    if (c->side == nullptr) {
        GP_CUDA_CHECK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
        GP_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        GP_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    GP_CUDA_CHECK(cudaEventRecord(c->ev_fork, stream));
    GP_CUDA_CHECK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    GP_LAUNCH(rowsort_big_kernel, gp_sm_count() * 4, BIG_THREADS, (size_t)c->big_smem_bytes, c->side, ptr, (const int *)len,
              colbuf, len, (const int *)c->biglist, c->meta, max_word, c->bitmap_words, (const int *)c->medlist);
    GP_CUDA_CHECK(cudaEventRecord(c->ev_join, c->side));
    GP_LAUNCH_PDL(rowsort_small_kernel, launch_blocks(n, 256), 256, 0, stream, ptr, (const int *)len, colbuf, (long long)n, len,
                  c->biglist, c->meta, max_word, 0);
    GP_CUDA_CHECK(cudaStreamWaitEvent(stream, c->ev_join, 0));
    return GP_OK;
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
    c->hub_capacity = kcap / (GP_CHUNK_EDGES + 1) + 1;
    {
        // region of class c: hub chunks (<= E/128 + one partial chunk per hub row), then rows of more than 64, 32, 16, 8,
        // 4 edges (a class cannot hold more rows than E / its smallest degree, nor more than N), then everything else
        const int64_t min_deg[GP_NUM_CLASSES] = {0, 65, 33, 17, 9, 5, 0};
        int64_t off = 0;
        for (int k = 0; k < GP_NUM_CLASSES; ++k) {
            int64_t cap = k == 0 ? kcap / GP_CHUNK_EDGES + c->hub_capacity + 1
                                 : (min_deg[k] ? std::min<int64_t>(num_nodes, kcap / min_deg[k]) : num_nodes) + 1;
            c->desc_off[k] = (int)off;
            off += cap;
        }
        GP_REQUIRE(off < (1ll << 31), GP_ERR_UNSUPPORTED, "gp_csr_create: work list too large for this build");
        c->desc_off[GP_NUM_CLASSES] = (int)off;
        c->desc_capacity = off;
    }
    c->big_capacity = kcap / 129 + 1;
    // node bitmap for the long-row sort, if it fits next to a few resident CTAs
    const int64_t words = (num_nodes + 31) / 32;
    c->bitmap_words = words * 4 <= 96 * 1024 ? (int)words : 0;
    c->big_smem_bytes = BIG_SMEM_ELEMS * (int)sizeof(int);
    if (c->bitmap_words * 4 > c->big_smem_bytes) c->big_smem_bytes = (c->bitmap_words * 4 + 15) / 16 * 16;
    if (cudaFuncSetAttribute(rowsort_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) {
        cudaGetLastError();
        c->bitmap_words = c->bitmap_words * 4 <= BIG_SMEM_ELEMS * (int)sizeof(int) ? c->bitmap_words : 0;
        c->big_smem_bytes = BIG_SMEM_ELEMS * (int)sizeof(int);
    }
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
    alloc((void **)&c->row_start, (nn + 1) * sizeof(int));
    alloc((void **)&c->cursor, (nn + 1) * sizeof(int));
    alloc((void **)&c->col, kc * sizeof(int));
    alloc((void **)&c->biglist, (size_t)c->big_capacity * sizeof(int));
    c->med_capacity = kcap / 17 + 1;
    alloc((void **)&c->medlist, (size_t)c->med_capacity * sizeof(int));
    alloc((void **)&c->desc, (size_t)c->desc_capacity * sizeof(int4));
    // meta words, then the look-back words of the two scans of a build (separate regions, so one
    // memset per build clears everything), then the ticket counters
    c->scan_b_offset = (size_t)scan_tiles(num_nodes + 1) * ScanStatus<1>::STRIDE;
    c->scan_status_words = c->scan_b_offset + (size_t)scan_tiles(num_nodes + 1) * ScanStatus<ROW_CH>::STRIDE + 8;
    // one allocation, cleared by ONE memset per build: meta | look-back words | raw edge counts (deg)
    alloc((void **)&c->meta, (GP_META_WORDS + c->scan_status_words + nn + 1) * sizeof(int));
    c->scan_status = c->meta != nullptr ? c->meta + GP_META_WORDS : nullptr;
    c->deg = c->meta != nullptr ? c->scan_status + c->scan_status_words : nullptr;
    if (rc != GP_OK) {
        gp_csr_free(c);
        return rc;
    }
    *out = c;
    return GP_OK;
}

int gp_csr_scratch(gp_csr *c, int slot, size_t bytes, void **out)
{
    GP_REQUIRE(c != nullptr && out != nullptr && slot >= 0 && slot < 16, GP_ERR_INVALID, "gp_csr_scratch: bad argument");
    if (bytes == 0) bytes = 16;
    if (c->scratch_bytes[slot] < bytes) {
        if (c->scratch[slot]) cudaFree(c->scratch[slot]);
        c->scratch[slot] = nullptr;
        c->scratch_bytes[slot] = 0;
        GP_CUDA_CHECK(cudaMalloc(&c->scratch[slot], bytes));
        c->scratch_bytes[slot] = bytes;
    }
    *out = c->scratch[slot];
    return GP_OK;
}

extern "C" int gp_csr_free(gp_csr_t *c)
{
    if (!c) return GP_OK;
    for (int i = 0; i < 16; ++i) cudaFree(c->scratch[i]);
    for (int i = 0; i < 12; ++i)
        if (c->trace_ev[i]) cudaEventDestroy(c->trace_ev[i]);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->side) cudaStreamDestroy(c->side);
    cudaFree(c->row_start);
    cudaFree(c->cursor);
    cudaFree(c->col);
    cudaFree(c->rowptr_in);
    cudaFree(c->col_in);
    cudaFree(c->deg_in);
    cudaFree(c->biglist);
    cudaFree(c->medlist);
    cudaFree(c->desc);
    cudaFree(c->meta);
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
    const int64_t n = c->num_nodes;
    c->num_input_edges = num_edges;
    c->built = false;
    c->in_built = false;
    c->trace_n = 0;
    auto mark = [&]() {  // GP_CSR_TRACE=1: an event after every launch (event nodes when captured)
        if (!gp_env().csr_trace || c->trace_n >= 12) return;
        if (c->trace_ev[c->trace_n] == nullptr) cudaEventCreate(&c->trace_ev[c->trace_n]);
        cudaEventRecordWithFlags(c->trace_ev[c->trace_n++], stream,
                                 gp_is_capturing() ? cudaEventRecordExternal : cudaEventRecordDefault);
    };
    mark();
    GP_CUDA_CHECK(cudaMemsetAsync(c->meta, 0, (GP_META_WORDS + c->scan_status_words + (size_t)n + 1) * sizeof(int), stream));
    if (n > 0) {
        const long long *ei = (const long long *)d_edge_index;
        const int vec = (reinterpret_cast<uintptr_t>(ei) & 15u) == 0 && num_edges % 2 == 0;  // both rows 16-byte aligned
        int *ticket_a = c->scan_status + c->scan_status_words - 8, *ticket_b = ticket_a + 1;
        mark();
        if (num_edges > 0)
            GP_LAUNCH(count_edges_kernel, launch_blocks(gp_ceil_div(num_edges, 4), 256), 256, 0, stream, ei, num_edges, n, sym,
                      vec, c->deg, c->meta);
        mark();
        PtrScanBuildIo pio{c->deg, c->row_start, c->cursor, n, c->biglist, c->medlist, c->meta};
        gp_count_launch();
        // (16 items per thread = 22 tiles = one look-back window at Flickr size measured SLOWER: 10.3 us against 8.8 us;
        // the per-thread serial work outweighs the saved look-back rounds)
        GP_CUDA_CHECK(gp_launch_pdl(chained_scan_kernel<1, PtrScanBuildIo, SCAN_IPT>, dim3((unsigned)scan_tiles(n + 1)),
                                    dim3(SCAN_THREADS), 0, stream, pio, (long long)(n + 1), c->scan_status, ticket_a));
        mark();
        if (num_edges > 0)
            GP_LAUNCH_PDL(scatter_edges_kernel, launch_blocks(gp_ceil_div(num_edges, 4), 256), 256, 0, stream, ei,
                          (long long)num_edges, (long long)n, sym, vec, c->cursor, c->col);
        mark();
        GP_TRY(sort_rows_forked(c, c->row_start, c->deg, c->col, GP_META_MAX_DEGREE, stream));  // deg := distinct degree
        mark();
        RowScanIo rio;  // writes the work-list descriptors as it goes
        rio.deg = c->deg;
        rio.row_start = c->row_start;
        rio.desc = c->desc;
        for (int k = 0; k < GP_NUM_CLASSES; ++k) rio.off[k] = c->desc_off[k];
        rio.meta = c->meta;
        rio.n = n;
        gp_count_launch();
        GP_CUDA_CHECK(gp_launch_pdl(chained_scan_kernel<ROW_CH, RowScanIo, SCAN_IPT>, dim3((unsigned)scan_tiles(n)),
                                    dim3(SCAN_THREADS), 0, stream, rio, (long long)n, c->scan_status + c->scan_b_offset,
                                    ticket_b));
        mark();
        mark();
    }
    GP_CUDA_CHECK(cudaGetLastError());
    c->built = true;
    return GP_OK;
}

// Diagnostics (GP_CSR_TRACE=1): milliseconds between the marks of the last build: memsets, count, scan, scatter,
// row sort (both kernels), row scan, descriptors.  syncs on the last mark.  Returns the number of intervals.
extern "C" int gp_csr_trace_ms(gp_csr_t *c, float *ms, int32_t cap, int32_t *num)
{
    GP_REQUIRE(c != nullptr && ms != nullptr && num != nullptr, GP_ERR_INVALID, "gp_csr_trace_ms: NULL argument");
    GP_REQUIRE(c->trace_n >= 2, GP_ERR_INVALID, "gp_csr_trace_ms: set GP_CSR_TRACE=1 and build first");
    GP_CUDA_CHECK(cudaEventSynchronize(c->trace_ev[c->trace_n - 1]));
    int k = 0;
    for (; k + 1 < c->trace_n && k < cap; ++k) GP_CUDA_CHECK(cudaEventElapsedTime(ms + k, c->trace_ev[k], c->trace_ev[k + 1]));
    *num = k;
    return GP_OK;
}

int gp_csr_ensure_in(gp_csr *c, cudaStream_t stream)
{
    GP_REQUIRE(c != nullptr && c->built, GP_ERR_INVALID, "in-edge CSR requested before gp_csr_build");
    if (c->in_built) return GP_OK;
    const int64_t n = c->num_nodes;
    if (c->rowptr_in == nullptr) {  // the transposed arrays are allocated on first use
        GP_REQUIRE(!gp_is_capturing(), GP_ERR_INVALID, "in-edge CSR cannot be allocated inside a graph capture");
        const size_t kc = (size_t)(c->key_capacity > 0 ? c->key_capacity : 1);
        GP_CUDA_CHECK(cudaMalloc((void **)&c->rowptr_in, (size_t)(n + 1) * sizeof(int)));
        GP_CUDA_CHECK(cudaMalloc((void **)&c->col_in, kc * sizeof(int)));
        GP_CUDA_CHECK(cudaMalloc((void **)&c->deg_in, (size_t)(n + 1) * sizeof(int)));
    }
    GP_LAUNCH(set_meta_kernel, 1, 1, 0, stream, c->meta, GP_META_IS_SYMMETRIC, 1);
    if (n > 0) {
        // cursor / scan_status are free again once the out-edge CSR is finished... except that cursor
        // is only the scatter cursor of the build, which has finished
        int *ticket = c->scan_status + c->scan_status_words - 8;
        GP_CUDA_CHECK(cudaMemsetAsync(c->deg_in, 0, (size_t)(n + 1) * sizeof(int), stream));
        GP_CUDA_CHECK(cudaMemsetAsync(c->scan_status, 0, c->scan_status_words * sizeof(int), stream));
        GP_LAUNCH(count_in_kernel, row_blocks(n), 256, 0, stream, c->row_start, c->deg, c->col, n, c->deg_in);
        PtrScanIo pio{c->deg_in, c->rowptr_in, c->cursor, n};
        gp_count_launch();
        chained_scan_kernel<1, PtrScanIo><<<scan_tiles(n + 1), SCAN_THREADS, 0, stream>>>(pio, n + 1, c->scan_status,
                                                                                        ticket);
        GP_LAUNCH(scatter_in_kernel, row_blocks(n), 256, 0, stream, c->row_start, c->deg, c->col, n, c->cursor,
                  c->col_in);
        // the transposed rows hold no duplicates: sorting them in place restores ascending order
        GP_TRY(sort_rows(c, c->rowptr_in, c->deg_in, c->col_in, GP_META_SCRATCH, true, stream));
        GP_LAUNCH(compare_csr_kernel, row_blocks(n), 256, 0, stream, c->row_start, c->deg, c->col, c->rowptr_in,
                  c->col_in, n, c->meta);
    }
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
    const int64_t n = c->num_nodes;
    if (n == 0) {
        GP_CUDA_CHECK(cudaMemsetAsync(d_rowptr, 0, sizeof(int), stream));
        GP_CUDA_CHECK(cudaStreamSynchronize(stream));
        return GP_OK;
    }
    if (which == 1) {
        GP_TRY(gp_csr_ensure_in(c, stream));
        int m = 0;
        GP_CUDA_CHECK(cudaMemcpyAsync(&m, &c->meta[GP_META_NUM_EDGES], sizeof(int), cudaMemcpyDeviceToHost, stream));
        GP_CUDA_CHECK(cudaStreamSynchronize(stream));
        GP_CUDA_CHECK(cudaMemcpyAsync(d_rowptr, c->rowptr_in, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToDevice,
                                      stream));
        if (m > 0)
            GP_CUDA_CHECK(cudaMemcpyAsync(d_col, c->col_in, (size_t)(u32)m * sizeof(int), cudaMemcpyDeviceToDevice,
                                          stream));
    } else {
        // squeeze the slack out: rowptr = exclusive prefix of deg, then a row-wise gather
        int *ticket = c->scan_status + c->scan_status_words - 8;
        GP_CUDA_CHECK(cudaMemsetAsync(c->scan_status, 0, c->scan_status_words * sizeof(int), stream));
        PtrScanIo pio{c->deg, d_rowptr, c->cursor, n};
        gp_count_launch();
        chained_scan_kernel<1, PtrScanIo><<<scan_tiles(n + 1), SCAN_THREADS, 0, stream>>>(pio, n + 1, c->scan_status,
                                                                                        ticket);
        GP_LAUNCH(gather_rows_kernel, row_blocks(n), 256, 0, stream, c->row_start, c->deg, c->col, d_rowptr, n, d_col);
        GP_CUDA_CHECK(cudaGetLastError());
    }
    GP_CUDA_CHECK(cudaStreamSynchronize(stream));
    return GP_OK;
}
