// gp_cdist.cu — GraphPOPE-node2vec block (reference utils.py:174-176): pairwise
// {cosine distance, cosine similarity, euclidean distance} between the node2vec table
// emb[N, D] and the anchor embeddings a[K, D], then MinMaxScaler per column.
//
// This is the one dense contraction of the path, so it runs on the 5th-gen tensor cores:
//   dot[i, j] = emb_i . a_j   via tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate in TMEM).
// fp32 inputs are split on the fly into bf16 hi + bf16 lo (x = hi + lo + O(2^-17 x)) and three
// MMAs are accumulated (hi*hi + hi*lo + lo*hi), which keeps the dot product within ~1e-6 relative of
// fp32 — plain bf16 or tf32 would miss the 1e-4 parity bar.  Row norms are exact fp32 sums.
// Euclidean distances are sqrt(max(xx - 2 dot + aa, 0)) like sklearn; where that expansion cancels
// (a node coinciding with a stochastic anchor, true distance 0) the entry is recomputed from the
// difference form in fp32.
//
// Tiling: one persistent CTA per SM; a tile = 128 nodes x (<= 256 anchors).  The anchor operand (hi
// and lo, K-major, 128-byte swizzle) stays resident in shared memory, node tiles are staged by four
// producer warps, one thread issues the MMAs into one of two 256-column TMEM accumulators, and four
// epilogue warps drain the other one: thread = node row out of TMEM, a 32 x 32 shared-memory
// transpose, then thread = anchor column for the per-column min / max and for 128-byte coalesced row
// stores.  The kernel is HBM bound by construction (AI ~ 43 flop/B at D = 128): 4*N*D bytes in,
// 4*N*K bytes out; the MinMaxScaler costs a second pass over the inputs, not over the outputs.
#include "gp_internal.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>

namespace {

constexpr int CD_TILE_M = 128;
constexpr int CD_MAX_N = 256;   // anchors per column tile (UMMA N)
constexpr int CD_KCHUNK = 64;   // bf16 elements per 128-byte swizzle row

struct CdistParams {
    const float *emb;
    const float *anc;
    long long n, k, d;
    int mode;
    float *out;
    long long ld_out, col_offset;
    float *colmin;  // [k] running per-column min (ordered-int encoded)
    float *colmax;
    unsigned long long *best;  // k-means assignment mode: [n] packed (squared distance bits << 32 | column), min-reduced
};

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

// K-major, 128-byte-swizzle shared memory matrix descriptor (see cute/arch/mma_sm100_desc.hpp):
// start address >> 4 in bits [0,14), leading byte offset (unused for swizzled K-major, 1) in
// [16,30), stride byte offset = 1024 B (8 rows x 128 B) >> 4 in [32,46), version 1 in [46,48),
// layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ u64 make_sw128_desc(u32 saddr)
{
    u64 d = 0;
    d |= (u64)((saddr >> 4) & 0x3FFFu);
    d |= (u64)1 << 16;
    d |= (u64)(1024 >> 4) << 32;
    d |= (u64)1 << 46;
    d |= (u64)2 << 61;
    return d;
}

// kind::f16 instruction descriptor: D = F32 (bits [4,6) = 1), A = B = BF16 (bits [7,10), [10,13) = 1),
// both K-major (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29).
__host__ __device__ constexpr u32 make_idesc(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(n >> 3) << 17) | ((u32)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity)
{
    u32 done;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(40);  // waiting warps share their scheduler with the warps they wait for
    }
}

// Orderable-int encoding so float min/max can use integer atomics for any sign.
__device__ __forceinline__ int f2ord(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__device__ __forceinline__ void mbar_init(u32 bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(u32 bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 8 fp32 values -> one 16-byte group of bf16 "hi" and one of bf16 "lo" (x = hi + lo + O(2^-17 x)),
// stored at 16-byte column c ^ (row & 7) of the row's 128-byte swizzle line; returns the sum of squares.
__device__ __forceinline__ float split_store8(const float4 a, const float4 b, unsigned char *s_hi, unsigned char *s_lo,
                                              size_t off)
{
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    u32 hi[4], lo[4];
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float x0 = x[2 * i], x1 = x[2 * i + 1];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
        hi[i] = (u32)__bfloat16_as_ushort(h0) | ((u32)__bfloat16_as_ushort(h1) << 16);
        lo[i] = (u32)__bfloat16_as_ushort(l0) | ((u32)__bfloat16_as_ushort(l1) << 16);
        ss = fmaf(x0, x0, ss);
        ss = fmaf(x1, x1, ss);
    }
    *reinterpret_cast<uint4 *>(s_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(s_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    return ss;
}

// Stage `rows_pad` fp32 rows of length D (row-major, leading dim D) as bf16 hi / lo operands in the
// canonical K-major SWIZZLE_128B layout: [D/64 chunks][rows_pad][128 B].  Rows >= valid are zero.
// Also writes the exact fp32 row norms.  `nthreads` consecutive threads starting at `t0` take part;
// a warp reads whole rows (32 bytes per lane), UNROLL row groups in flight per thread.
template <int D, int UNROLL>
__device__ __forceinline__ void stage_operand(const float *__restrict__ src, long long first_row, long long valid,
                                              int rows_pad, unsigned char *s_hi, unsigned char *s_lo, float *s_norm,
                                              int t, int nthreads)
{
    constexpr int GROUPS = D / 8;  // 8-element (16-byte bf16) groups per row
    const int total = rows_pad * GROUPS;
    for (int g0 = t; g0 < total; g0 += nthreads * UNROLL) {
        float4 a[UNROLL], b[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int g = g0 + u * nthreads;
            const int row = g / GROUPS, kg = g % GROUPS;
            a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            b[u] = a[u];
            if (g < total && (long long)row < valid) {
                const float4 *q = reinterpret_cast<const float4 *>(src + (size_t)(first_row + row) * D + kg * 8);
                a[u] = __ldg(q);
                b[u] = __ldg(q + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int g = g0 + u * nthreads;
            const int row = g / GROUPS, kg = g % GROUPS;
            const int chunk = kg / 8, c = kg % 8;
            const size_t off = (size_t)chunk * rows_pad * 128 + (size_t)row * 128 + (size_t)((c ^ (row & 7)) * 16);
            float ss = 0.0f;
            if (g < total) ss = split_store8(a[u], b[u], s_hi, s_lo, off);
            // row norm: the GROUPS threads of a row are consecutive lanes (GROUPS = 8 or 16 divides 32)
#pragma unroll
            for (int m = GROUPS / 2; m; m >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, m);
            if (g < total && kg == 0) s_norm[row] = ss;
        }
    }
}

// Euclidean entries where xx - 2 dot + aa cancels (a node coinciding with an anchor): recomputed from the
// differences in fp32.  Kept out of line: it is rare, and the epilogue loop must stay small enough
// for the instruction cache (the first version of this kernel was 160 KB of SASS and spent most of
// its time waiting for instruction fetches).
template <int D>
__device__ __noinline__ float diff_form_d2(const float *__restrict__ xr, const float *__restrict__ ar)
{
    float s = 0.0f;
    for (int t = 0; t < D; ++t) {
        const float df = __ldg(xr + t) - __ldg(ar + t);
        s = fmaf(df, df, s);
    }
    return s;
}

// Warp roles: warps 0-7 epilogue (TMEM lane quadrant = warp id % 4, column half = warp id / 4: a lone
// warp per scheduler cannot hide its own ALU latency), warps 8-11 operand producers, warp 12 issues
// the MMAs.  Per 128-row tile: producers stage A (hi, lo) once the previous tile's
// MMAs have released the buffer; the MMA warp accumulates hi*hi + hi*lo + lo*hi into one of two
// TMEM accumulators; the epilogue of tile i overlaps the staging and MMAs of tile i + 1.
constexpr int CD_EPI_WARPS = 8;
constexpr int CD_PROD_WARPS = 4;
constexpr int CD_THREADS2 = 32 * (CD_EPI_WARPS + CD_PROD_WARPS + 1);
constexpr int CD_TR_LD = 17;  // padded leading dimension of the 32 x 16 transpose tiles

template <int D>
__global__ void __launch_bounds__(CD_THREADS2, 1) cdist_kernel(CdistParams p, int n_pad, int write_out, int apply_scale,
                                                               int track_minmax)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: B hi | B lo | A hi | A lo (each 1024-aligned), then transpose tiles, norms, scales, barriers
    // (an offset, not an integer round trip of the pointer: the compiler must keep seeing shared memory,
    // or every access below turns into a generic LD / ST)
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const size_t b_bytes = (size_t)(D / CD_KCHUNK) * n_pad * 128;
    const size_t a_bytes = (size_t)(D / CD_KCHUNK) * CD_TILE_M * 128;
    unsigned char *sB_hi = smem, *sB_lo = sB_hi + b_bytes;
    unsigned char *sA_hi = sB_lo + b_bytes, *sA_lo = sA_hi + a_bytes;
    float *s_tr = reinterpret_cast<float *>(sA_lo + a_bytes);       // [8 warps][32][CD_TR_LD]
    float *s_an = s_tr + CD_EPI_WARPS * 32 * CD_TR_LD;              // [CD_MAX_N] anchor norms
    float *s_xn = s_an + CD_MAX_N;                                  // [3][CD_TILE_M] node norms, by tile % 3
    float *s_scale = s_xn + 3 * CD_TILE_M;                          // [CD_MAX_N]
    float *s_shift = s_scale + CD_MAX_N;                            // [CD_MAX_N]
    float *s_ian = s_shift + CD_MAX_N;                              // [CD_MAX_N] 1 / |anchor| (0 for a zero row)
    float *s_cmm = s_ian + CD_MAX_N;                                // [8 warps][2][CD_MAX_N / 2] running column min / max
    u64 *s_bar = reinterpret_cast<u64 *>(s_cmm + CD_EPI_WARPS * CD_MAX_N);  // a_full, a_empty, acc_full[2], acc_empty[2]
    u32 *s_tmem = reinterpret_cast<u32 *>(s_bar + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long col_tile = blockIdx.y;
    const long long k0 = col_tile * CD_MAX_N;
    const long long kvalid = min((long long)CD_MAX_N, p.k - k0);
    const u32 bar_a_full = smem_u32(s_bar + 0), bar_a_empty = smem_u32(s_bar + 1);
    const u32 bar_acc_full = smem_u32(s_bar + 2), bar_acc_empty = smem_u32(s_bar + 4);  // [2] each, 8 bytes apart

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar_a_full, CD_PROD_WARPS * 32);
        mbar_init(bar_a_empty, 1);
        mbar_init(bar_acc_full, 1);
        mbar_init(bar_acc_full + 8, 1);
        mbar_init(bar_acc_empty, CD_EPI_WARPS * 32);
        mbar_init(bar_acc_empty + 8, CD_EPI_WARPS * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // anchors of this column tile: resident for the CTA's lifetime (all threads help)
    stage_operand<D, 2>(p.anc, k0, kvalid, n_pad, sB_hi, sB_lo, s_an, tid, CD_THREADS2 / 32 * 32);
    for (int c = tid; c < n_pad; c += CD_THREADS2) {
        float scale = 1.0f, shift = 0.0f;
        if (apply_scale && c < kvalid) {
            // MinMaxScaler.transform (utils.py:176): X * scale + min_, scale = 1/range (range < 10 eps -> 1)
            const float dmin = ord2f(reinterpret_cast<const int *>(p.colmin)[k0 + c]);
            const float dmax = ord2f(reinterpret_cast<const int *>(p.colmax)[k0 + c]);
            float range = dmax - dmin;
            if (range < 10.0f * 1.1920929e-07f) range = 1.0f;  // sklearn _handle_zeros_in_scale
            scale = __fdiv_rn(1.0f, range);
            shift = __fsub_rn(0.0f, __fmul_rn(dmin, scale));
        }
        s_scale[c] = scale;
        s_shift[c] = shift;
    }
    __syncthreads();  // s_an complete
    for (int c = tid; c < n_pad; c += CD_THREADS2) s_ian[c] = s_an[c] > 0.0f ? rsqrtf(s_an[c]) : 0.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_base = *s_tmem;
    const long long row_tiles = (p.n + CD_TILE_M - 1) / CD_TILE_M;

    if (warp >= CD_EPI_WARPS && warp < CD_EPI_WARPS + CD_PROD_WARPS) {
        // ------------------------------------------------------------ producers
        const int pt = tid - CD_EPI_WARPS * 32;
        long long it = 0;
        for (long long tile = blockIdx.x; tile < row_tiles; tile += gridDim.x, ++it) {
            const long long r0 = tile * CD_TILE_M;
            const long long rvalid = min((long long)CD_TILE_M, p.n - r0);
            if (it > 0) mbar_wait(bar_a_empty, (u32)((it - 1) & 1));  // the MMAs of the previous tile have read A
            stage_operand<D, 8>(p.emb, r0, rvalid, CD_TILE_M, sA_hi, sA_lo, s_xn + (it % 3) * CD_TILE_M, pt,
                                CD_PROD_WARPS * 32);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor core reads
            mbar_arrive(bar_a_full);
        }
    } else if (warp == CD_EPI_WARPS + CD_PROD_WARPS) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const u32 idesc = make_idesc(CD_TILE_M, n_pad);
            long long it = 0;
            for (long long tile = blockIdx.x; tile < row_tiles; tile += gridDim.x, ++it) {
                const int st = (int)(it & 1);
                const long long use = it >> 1;  // earlier uses of this accumulator
                mbar_wait(bar_a_full, (u32)(it & 1));
                if (use > 0) mbar_wait(bar_acc_empty + 8 * st, (u32)((use - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const u32 tmem_d = tmem_base + (u32)(st * CD_MAX_N);
                u32 acc = 0;
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const unsigned char *a_src = term == 2 ? sA_lo : sA_hi;  // hi*hi, hi*lo, lo*hi
                    const unsigned char *b_src = term == 1 ? sB_lo : sB_hi;
#pragma unroll
                    for (int kc = 0; kc < D / CD_KCHUNK; ++kc) {
#pragma unroll
                        for (int ks = 0; ks < CD_KCHUNK / 16; ++ks) {
                            const u64 da = make_sw128_desc(smem_u32(a_src + (size_t)kc * CD_TILE_M * 128 + ks * 32));
                            const u64 db = make_sw128_desc(smem_u32(b_src + (size_t)kc * n_pad * 128 + ks * 32));
                            umma_bf16(tmem_d, da, db, idesc, acc);
                            acc = 1;
                        }
                    }
                }
                umma_commit(bar_a_empty);            // A may be overwritten
                umma_commit(bar_acc_full + 8 * st);  // accumulator ready
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue
        const int quad = warp & 3, half = warp >> 2;         // TMEM lane quadrant, column half
        const int ncols = n_pad >> 1, cbase = half * ncols;  // this warp's columns [cbase, cbase + ncols)
        float *tr = s_tr + warp * 32 * CD_TR_LD;
        float *wmin = s_cmm + warp * CD_MAX_N, *wmax = wmin + CD_MAX_N / 2;  // running min / max of its columns
        for (int c = lane; c < CD_MAX_N / 2; c += 32) {
            wmin[c] = INFINITY;
            wmax[c] = -INFINITY;
        }
        __syncwarp();
        const int mode = p.mode;
        const int kv = (int)kvalid;
        const int rsel = lane >> 4, cl = lane & 15;  // transposed phase: two rows x 16 columns per step
        long long it = 0;
        for (long long tile = blockIdx.x; tile < row_tiles; tile += gridDim.x, ++it) {
            const int st = (int)(it & 1);
            const long long use = it >> 1;
            const long long r0 = tile * CD_TILE_M + quad * 32;
            mbar_wait(bar_acc_full + 8 * st, (u32)(use & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long row = r0 + lane;
            const bool row_ok = row < p.n;
            const float xx = s_xn[(it % 3) * CD_TILE_M + quad * 32 + lane];
            const float inv_xn = xx > 0.0f ? rsqrtf(xx) : 0.0f;
            const int rows_here = (int)min(32ll, max(0ll, p.n - r0));
            float best_d = INFINITY;
            int best_c = 0;
#pragma unroll 1
            for (int c0 = cbase; c0 < cbase + ncols; c0 += 16) {
                u32 v[16];
                const u32 taddr = tmem_base + ((u32)(quad * 32) << 16) + (u32)(st * CD_MAX_N + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                    "%14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (p.best != nullptr) {
                    // k-means assignment: nearest centre of this row among the 16 columns (squared distance
                    // xx - 2 dot + cc; no output block at all)
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float d2 = fmaxf(fmaf(-2.0f, __uint_as_float(v[j]), xx + s_an[c0 + j]), 0.0f);
                        if (c0 + j < kv && d2 < best_d) {
                            best_d = d2;
                            best_c = c0 + j;
                        }
                    }
                    continue;
                }
                // thread = node row: finish the 16 entries of this row, park them in the transpose tile
                if (mode == GP_CDIST_EUCLIDEAN) {
                    float d2[16];
                    u32 fix = 0;  // entries in the cancellation zone (rare): recomputed below, out of the main loop
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float sum = xx + s_an[c0 + j];
                        d2[j] = fmaf(-2.0f, __uint_as_float(v[j]), sum);
                        fix |= (d2[j] < 1.0e-3f * sum ? 1u : 0u) << j;
                    }
                    if (!row_ok) fix = 0;
                    while (fix) {
                        const int j = __ffs(fix) - 1;
                        fix &= fix - 1;
                        if (c0 + j < kv) {
                            const float e = diff_form_d2<D>(p.emb + (size_t)row * D, p.anc + (size_t)(k0 + c0 + j) * D);
#pragma unroll
                            for (int t = 0; t < 16; ++t)
                                if (t == j) d2[t] = e;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float r;
                        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(d2[j], 0.0f)));
                        tr[lane * CD_TR_LD + j] = r;
                    }
                } else {
                    const bool sim_mode = mode == GP_CDIST_COSINE_SIMILARITY;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float sim = __uint_as_float(v[j]) * inv_xn * s_ian[c0 + j];
                        tr[lane * CD_TR_LD + j] = sim_mode ? sim : fminf(fmaxf(1.0f - sim, 0.0f), 2.0f);
                    }
                }
                __syncwarp();
                // lanes 0-15 / 16-31 = the 16 anchor columns of an even / odd row: running min / max,
                // scaling, and row stores of 64 contiguous bytes
                const int c = c0 + cl;
                const bool cvalid = c < kv;
                const float sc = s_scale[c], sh = s_shift[c];
                float *ocol = p.out + (size_t)(r0 + rsel) * p.ld_out + p.col_offset + k0 + c;
                float mn = INFINITY, mx = -INFINITY;
                if (write_out && cvalid) {
                    if (apply_scale) {
#pragma unroll 4
                        for (int rr = rsel; rr < rows_here; rr += 2)
                            ocol[(size_t)(rr - rsel) * p.ld_out] = __fadd_rn(__fmul_rn(tr[rr * CD_TR_LD + cl], sc), sh);
                    } else {
#pragma unroll 4
                        for (int rr = rsel; rr < rows_here; rr += 2) ocol[(size_t)(rr - rsel) * p.ld_out] = tr[rr * CD_TR_LD + cl];
                    }
                }
                if (track_minmax) {
#pragma unroll 4
                    for (int rr = rsel; rr < rows_here; rr += 2) {
                        const float val = tr[rr * CD_TR_LD + cl];
                        mn = fminf(mn, val);
                        mx = fmaxf(mx, val);
                    }
                    mn = fminf(mn, __shfl_xor_sync(FULL_MASK, mn, 16));
                    mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, 16));
                    if (lane < 16) {
                        wmin[c - cbase] = fminf(wmin[c - cbase], mn);
                        wmax[c - cbase] = fmaxf(wmax[c - cbase], mx);
                    }
                }
                __syncwarp();
            }
            if (p.best != nullptr && row_ok && best_d < INFINITY)
                atomicMin(p.best + row, ((unsigned long long)__float_as_uint(best_d) << 32) | (unsigned long long)(u32)(k0 + best_c));
            // this warp has drained its part of the accumulator
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_acc_empty + 8 * st);
        }
        if (track_minmax) {
            for (int c = lane; c < ncols; c += 32) {
                if (cbase + c < kvalid && wmin[c] <= wmax[c]) {
                    atomicMin(reinterpret_cast<int *>(p.colmin) + k0 + cbase + c, f2ord(wmin[c]));
                    atomicMax(reinterpret_cast<int *>(p.colmax) + k0 + cbase + c, f2ord(wmax[c]));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

__global__ void minmax_init_kernel(int *cmin, int *cmax, long long k)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < k) {
        cmin[c] = f2ord(INFINITY);
        cmax[c] = f2ord(-INFINITY);
    }
}

// ---------------------------------------------------------------- any embedding width
// The tensor-core kernel stages operands in 64-element K chunks and is instantiated for D = 64 and 128 (the
// reference writes 128-d tables, generate_node2vec_embedding.py:23).  utils.py:174 accepts any width, so:
//   D < 128, D not 64  rows are zero-padded to 64 / 128 in stream-ordered scratch (zeros change neither dot products
//                      nor norms) and take the tensor-core path;
//   D > 128            a plain fp32 kernel (shared-memory tiles, difference form for euclidean: no cancellation).
__global__ void pad_rows_kernel(const float *__restrict__ src, long long rows, int d, int dp, float *__restrict__ dst)
{
    const long long total = rows * dp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / dp;
        const int c = (int)(i - r * dp);
        dst[i] = c < d ? src[r * d + c] : 0.0f;
    }
}

constexpr int CG_TILE = 64, CG_KC = 16;

__global__ void __launch_bounds__(256)
cdist_generic_kernel(CdistParams p, int write_out, int track_minmax)
{
    __shared__ float s_x[CG_KC][CG_TILE + 1], s_a[CG_KC][CG_TILE + 1];
    __shared__ int s_min[CG_TILE], s_max[CG_TILE];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // thread = 4 rows (ty) x 4 columns (tx)
    const long long r0 = (long long)blockIdx.x * CG_TILE, c0 = (long long)blockIdx.y * CG_TILE;
    if (threadIdx.x < CG_TILE) {
        s_min[threadIdx.x] = f2ord(INFINITY);
        s_max[threadIdx.x] = f2ord(-INFINITY);
    }
    float acc[4][4], xx[4], aa[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        xx[i] = aa[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    }
    const bool euclid = p.mode == GP_CDIST_EUCLIDEAN;
    for (long long k0 = 0; k0 < p.d; k0 += CG_KC) {
        __syncthreads();
        for (int i = threadIdx.x; i < CG_TILE * CG_KC; i += 256) {
            const int rr = i / CG_KC, kk = i % CG_KC;
            const bool kin = k0 + kk < p.d;
            s_x[kk][rr] = (kin && r0 + rr < p.n) ? p.emb[(size_t)(r0 + rr) * p.d + k0 + kk] : 0.0f;
            s_a[kk][rr] = (kin && c0 + rr < p.k) ? p.anc[(size_t)(c0 + rr) * p.d + k0 + kk] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CG_KC; ++kk) {
            float xv[4], av[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xv[i] = s_x[kk][ty * 4 + i];
                av[i] = s_a[kk][tx * 4 + i];
                xx[i] = fmaf(xv[i], xv[i], xx[i]);
                aa[i] = fmaf(av[i], av[i], aa[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (euclid) {
                        const float df = xv[i] - av[j];
                        acc[i][j] = fmaf(df, df, acc[i][j]);
                    } else {
                        acc[i][j] = fmaf(xv[i], av[j], acc[i][j]);
                    }
                }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long c = c0 + tx * 4 + j;
        float mn = INFINITY, mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = r0 + ty * 4 + i;
            float v;
            if (euclid) {
                v = sqrtf(acc[i][j]);
            } else {
                const float ix = xx[i] > 0.0f ? rsqrtf(xx[i]) : 0.0f, ia = aa[j] > 0.0f ? rsqrtf(aa[j]) : 0.0f;
                const float sim = acc[i][j] * ix * ia;
                v = p.mode == GP_CDIST_COSINE_SIMILARITY ? sim : fminf(fmaxf(1.0f - sim, 0.0f), 2.0f);
            }
            if (r < p.n && c < p.k) {
                if (write_out) p.out[(size_t)r * p.ld_out + p.col_offset + c] = v;
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
            }
        }
        if (track_minmax && mn <= mx) {
            atomicMin(&s_min[tx * 4 + j], f2ord(mn));
            atomicMax(&s_max[tx * 4 + j], f2ord(mx));
        }
    }
    if (track_minmax) {
        __syncthreads();
        if (threadIdx.x < CG_TILE && c0 + threadIdx.x < p.k && s_min[threadIdx.x] <= s_max[threadIdx.x]) {
            atomicMin(reinterpret_cast<int *>(p.colmin) + c0 + threadIdx.x, s_min[threadIdx.x]);
            atomicMax(reinterpret_cast<int *>(p.colmax) + c0 + threadIdx.x, s_max[threadIdx.x]);
        }
    }
}

// MinMaxScaler.transform in place (utils.py:176), same arithmetic as the fused path of cdist_kernel.
__global__ void minmax_scale_kernel(CdistParams p)
{
    const long long total = p.n * p.k;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / p.k, c = i - r * p.k;
        const float dmin = ord2f(reinterpret_cast<const int *>(p.colmin)[c]);
        const float dmax = ord2f(reinterpret_cast<const int *>(p.colmax)[c]);
        float range = dmax - dmin;
        if (range < 10.0f * 1.1920929e-07f) range = 1.0f;
        const float scale = __fdiv_rn(1.0f, range), shift = __fsub_rn(0.0f, __fmul_rn(dmin, scale));
        float *o = p.out + (size_t)r * p.ld_out + p.col_offset + c;
        *o = __fadd_rn(__fmul_rn(*o, scale), shift);
    }
}

cudaMemPool_t cdist_pool()
{
    // stream-ordered scratch from a pool of our own that keeps its memory across synchronisations
    // (the default pool hands it back to the driver at every sync, ~100 us per call)
    static cudaMemPool_t pool = nullptr;
    if (pool == nullptr) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) return nullptr;
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return pool;
}

size_t cdist_smem_bytes(int d, int n_pad)
{
    const size_t b = (size_t)(d / CD_KCHUNK) * n_pad * 128, a = (size_t)(d / CD_KCHUNK) * CD_TILE_M * 128;
    return 1024 + 2 * b + 2 * a + sizeof(float) * (CD_EPI_WARPS * 32 * CD_TR_LD + 4 * CD_MAX_N + 3 * CD_TILE_M + CD_EPI_WARPS * CD_MAX_N) + 128;
}

}  // namespace

extern "C" int gp_cdist_minmax(const float *d_emb, const float *d_anchor_emb, int64_t num_nodes,
                               int64_t num_anchors, int64_t dim, int32_t mode, int32_t apply_minmax, float *d_out,
                               int64_t ld_out, int64_t col_offset, gp_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(num_nodes >= 0 && num_anchors >= 0 && dim > 0, GP_ERR_INVALID, "gp_cdist_minmax: bad sizes");
    GP_REQUIRE(mode >= 0 && mode <= 2, GP_ERR_INVALID, "gp_cdist_minmax: unknown mode %d", mode);
    GP_REQUIRE(col_offset >= 0 && ld_out >= col_offset + num_anchors, GP_ERR_INVALID,
               "gp_cdist_minmax: ld_out too small");
    if (num_nodes == 0 || num_anchors == 0) return GP_OK;
    GP_REQUIRE(d_emb != nullptr && d_anchor_emb != nullptr && d_out != nullptr, GP_ERR_INVALID,
               "gp_cdist_minmax: NULL argument");
    cudaMemPool_t pool = cdist_pool();
    GP_REQUIRE(pool != nullptr, GP_ERR_CUDA, "gp_cdist_minmax: cannot create the scratch memory pool");
    if (dim != 64 && dim != 128 && dim < 128) {
        // zero-pad the rows to the next instantiated width and take the tensor-core path
        const int dp = dim < 64 ? 64 : 128;
        float *emb_p = nullptr, *anc_p = nullptr;
        GP_CUDA_CHECK(cudaMallocFromPoolAsync((void **)&emb_p, sizeof(float) * (size_t)num_nodes * dp, pool, stream));
        GP_CUDA_CHECK(cudaMallocFromPoolAsync((void **)&anc_p, sizeof(float) * (size_t)num_anchors * dp, pool, stream));
        GP_LAUNCH(pad_rows_kernel, gp_sm_count() * 8, 256, 0, stream, d_emb, (long long)num_nodes, (int)dim, dp, emb_p);
        GP_LAUNCH(pad_rows_kernel, gp_sm_count() * 8, 256, 0, stream, d_anchor_emb, (long long)num_anchors, (int)dim, dp, anc_p);
        const int rc = gp_cdist_minmax(emb_p, anc_p, num_nodes, num_anchors, dp, mode, apply_minmax, d_out, ld_out,
                                       col_offset, stream_);
        GP_CUDA_CHECK(cudaFreeAsync(emb_p, stream));
        GP_CUDA_CHECK(cudaFreeAsync(anc_p, stream));
        return rc;
    }
    if (dim > 128) {
        CdistParams g;
        g.emb = d_emb;
        g.anc = d_anchor_emb;
        g.n = num_nodes;
        g.k = num_anchors;
        g.d = dim;
        g.mode = mode;
        g.out = d_out;
        g.ld_out = ld_out;
        g.col_offset = col_offset;
        g.colmin = g.colmax = nullptr;
        g.best = nullptr;
        int *mm = nullptr;
        if (apply_minmax) {
            GP_CUDA_CHECK(cudaMallocFromPoolAsync((void **)&mm, sizeof(int) * 2 * (size_t)num_anchors, pool, stream));
            g.colmin = reinterpret_cast<float *>(mm);
            g.colmax = reinterpret_cast<float *>(mm + num_anchors);
            GP_LAUNCH(minmax_init_kernel, (unsigned)gp_ceil_div(num_anchors, 256), 256, 0, stream, mm, mm + num_anchors,
                      num_anchors);
        }
        const dim3 ggrid((unsigned)gp_ceil_div(num_nodes, CG_TILE), (unsigned)gp_ceil_div(num_anchors, CG_TILE));
        GP_LAUNCH(cdist_generic_kernel, ggrid, 256, 0, stream, g, 1, apply_minmax ? 1 : 0);
        if (apply_minmax) {
            GP_LAUNCH(minmax_scale_kernel, gp_sm_count() * 8, 256, 0, stream, g);
            GP_CUDA_CHECK(cudaFreeAsync(mm, stream));
        }
        GP_CUDA_CHECK(cudaGetLastError());
        return GP_OK;
    }
    CdistParams p;
    p.emb = d_emb;
    p.anc = d_anchor_emb;
    p.n = num_nodes;
    p.k = num_anchors;
    p.d = dim;
    p.mode = mode;
    p.out = d_out;
    p.ld_out = ld_out;
    p.col_offset = col_offset;
    p.colmin = p.colmax = nullptr;
    p.best = nullptr;
    const int64_t col_tiles = gp_ceil_div(num_anchors, CD_MAX_N);
    const int64_t ktile = num_anchors < CD_MAX_N ? num_anchors : CD_MAX_N;
    const int n_pad = (int)(gp_ceil_div(ktile, 32) * 32);  // UMMA N (multiple of 16 at M = 128); the epilogue works in 32s
    const size_t smem = cdist_smem_bytes((int)dim, n_pad);
    const int64_t row_tiles = gp_ceil_div(num_nodes, CD_TILE_M);
    int grid_x = gp_sm_count();
    if (row_tiles < grid_x) grid_x = (int)row_tiles;
    const dim3 grid(grid_x, (unsigned)col_tiles);
    auto launch = [&](int write_out, int apply_scale, int track) -> int {
        gp_count_launch();
        if (dim == 128) {
            GP_CUDA_CHECK(cudaFuncSetAttribute(cdist_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cdist_kernel<128><<<grid, CD_THREADS2, smem, stream>>>(p, n_pad, write_out, apply_scale, track);
        } else {
            GP_CUDA_CHECK(cudaFuncSetAttribute(cdist_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cdist_kernel<64><<<grid, CD_THREADS2, smem, stream>>>(p, n_pad, write_out, apply_scale, track);
        }
        GP_CUDA_CHECK(cudaGetLastError());
        return GP_OK;
    };
    if (!apply_minmax) {
        GP_TRY(launch(1, 0, 0));
    } else {
        // MinMaxScaler (utils.py:175-176) without a second trip of the N x K block through HBM: pass 1
        // computes the block and keeps only the per-column min / max, pass 2 recomputes it (the node
        // table is L2 resident by then) and writes the scaled values.
        int *mm = nullptr;
        GP_CUDA_CHECK(cudaMallocFromPoolAsync((void **)&mm, sizeof(int) * 2 * (size_t)num_anchors, pool, stream));
        p.colmin = reinterpret_cast<float *>(mm);
        p.colmax = reinterpret_cast<float *>(mm + num_anchors);
        GP_LAUNCH(minmax_init_kernel, (unsigned)gp_ceil_div(num_anchors, 256), 256, 0, stream, mm, mm + num_anchors,
                  num_anchors);
        int rc = launch(0, 0, 1);
        if (rc == GP_OK) rc = launch(1, 1, 0);
        GP_CUDA_CHECK(cudaFreeAsync(mm, stream));
        GP_TRY(rc);
    }
    return GP_OK;
}


// ================================================================= KMeans on the device
// attach_node2vec with any sampling_method but 'stochastic' takes the KMeans cluster centres of the
// node2vec table as anchors (utils.py:168-170).  Lloyd iterations reuse the tcgen05 kernel above in
// "assignment" mode (nearest centre = arg-min over the pairwise block, which is never written out);
// k-means++ seeding draws the next centre with the exponential-race form of D^2 sampling (arg-min of
// Exp(1) / D^2), which needs a reduction instead of a prefix sum.  Parity with scikit-learn is
// statistical (unseeded there, utils.py:169): the tests compare inertia.
namespace {

__device__ __forceinline__ u64 splitmix64(u64 x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void fill_u64_kernel(unsigned long long *p, long long n, unsigned long long v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

constexpr int KPP_MAX_TRIALS = 8;

// One k-means++ step, first half: centre `step` := row chosen[step]; D^2 of every row updated with it;
// `trials` candidate rows for the next centre are drawn, each an exact D^2-weighted draw (arg-min of
// Exp(1) / D^2 over the rows, one independent race per candidate).  One warp per row.
template <int D>
__global__ void __launch_bounds__(256)
kpp_step_kernel(const float *__restrict__ x, long long n, const long long *chosen, int step, int k, int trials, u64 seed,
                float *__restrict__ mind2, float *__restrict__ centers, unsigned long long *next_key)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long c = chosen[step];
    constexpr int Q = D / 64;  // float2 per lane
    const float2 *crow = reinterpret_cast<const float2 *>(x + (size_t)c * D);
    float2 cv[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) cv[q] = __ldg(crow + q * 32 + lane);
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < Q; ++q) reinterpret_cast<float2 *>(centers + (size_t)step * D)[q * 32 + lane] = cv[q];
    }
    unsigned long long best[KPP_MAX_TRIALS];
#pragma unroll
    for (int l = 0; l < KPP_MAX_TRIALS; ++l) best[l] = ~0ull;
    for (long long i = warp; i < n; i += nwarps) {
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(x + (size_t)i * D) + q * 32 + lane);
            const float a = v.x - cv[q].x, b = v.y - cv[q].y;
            s += a * a + b * b;
        }
#pragma unroll
        for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(FULL_MASK, s, m);
        const float old = step == 0 ? INFINITY : mind2[i];
        const float d2 = fminf(old, s);
        if (lane == 0) mind2[i] = d2;
        if (d2 > 0.0f && step + 1 < k && lane < trials) {
            // lane l runs race l for this row
            const u64 r = splitmix64(seed ^ splitmix64(((u64)i * KPP_MAX_TRIALS + (u64)lane) * 0x100000001B3ull + (u64)step));
            const float u = ((float)(r >> 40) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
            const float key = -__logf(u) / d2;
            const unsigned long long packed = ((unsigned long long)__float_as_uint(key) << 32) | (unsigned long long)(u32)i;
#pragma unroll
            for (int l = 0; l < KPP_MAX_TRIALS; ++l)
                if (l == lane) best[l] = packed < best[l] ? packed : best[l];
        }
    }
#pragma unroll
    for (int l = 0; l < KPP_MAX_TRIALS; ++l)
        if (l == lane && lane < trials && best[l] != ~0ull) atomicMin(next_key + l, best[l]);
}

// Second half (greedy k-means++, as scikit-learn's 2 + log k local trials): the potential
// sum_i min(D^2_i, |x_i - candidate|^2) of every candidate.  One warp per row, candidates in shared memory.
template <int D>
__global__ void __launch_bounds__(256)
kpp_potential_kernel(const float *__restrict__ x, long long n, const unsigned long long *__restrict__ next_key, int trials,
                     const float *__restrict__ mind2, double *pot)
{
    __shared__ float s_cand[KPP_MAX_TRIALS][D];
    __shared__ double s_pot[KPP_MAX_TRIALS];
    const int lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < trials * D; t += blockDim.x) {
        const unsigned long long kx = next_key[t / D];
        const long long row = kx == ~0ull ? 0 : (long long)(kx & 0xFFFFFFFFull);
        s_cand[t / D][t % D] = __ldg(x + (size_t)row * D + (t % D));
    }
    if (threadIdx.x < KPP_MAX_TRIALS) s_pot[threadIdx.x] = 0.0;
    __syncthreads();
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    constexpr int Q = D / 32;
    double local[KPP_MAX_TRIALS];
#pragma unroll
    for (int l = 0; l < KPP_MAX_TRIALS; ++l) local[l] = 0.0;
    for (long long i = warp; i < n; i += nwarps) {
        float v[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) v[q] = __ldg(x + (size_t)i * D + q * 32 + lane);
        const float m = mind2[i];
#pragma unroll
        for (int l = 0; l < KPP_MAX_TRIALS; ++l) {
            if (l < trials) {
                float s = 0.0f;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float a = v[q] - s_cand[l][q * 32 + lane];
                    s += a * a;
                }
#pragma unroll
                for (int mm = 16; mm; mm >>= 1) s += __shfl_xor_sync(FULL_MASK, s, mm);
                local[l] += (double)fminf(m, s);
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int l = 0; l < KPP_MAX_TRIALS; ++l)
            if (l < trials) atomicAdd(&s_pot[l], local[l]);
    }
    __syncthreads();
    if (threadIdx.x < trials) atomicAdd(pot + threadIdx.x, s_pot[threadIdx.x]);
}

__global__ void kpp_pick_kernel(const unsigned long long *next_key, const double *pot, int trials, long long *chosen,
                                int step)
{
    int bl = -1;
    double bp = 0.0;
    for (int l = 0; l < trials; ++l) {
        if (next_key[l] == ~0ull) continue;
        if (bl < 0 || pot[l] < bp) {
            bl = l;
            bp = pot[l];
        }
    }
    // every row coincides with a centre already (fewer distinct rows than centres): reuse row 0
    chosen[step + 1] = bl < 0 ? 0 : (long long)(next_key[bl] & 0xFFFFFFFFull);
}

// sums[label] += row, counts[label]++, inertia += D^2 (one warp per row).
template <int D>
__global__ void __launch_bounds__(256)
kmeans_accumulate_kernel(const float *__restrict__ x, long long n, const unsigned long long *__restrict__ best,
                         float *sums, int *counts, double *inertia)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    double local = 0.0;
    for (long long i = warp; i < n; i += nwarps) {
        const unsigned long long kx = best[i];
        const int lab = (int)(kx & 0xFFFFFFFFull);
        float *dst = sums + (size_t)lab * D;
        for (int t = lane; t < D; t += 32) atomicAdd(dst + t, __ldg(x + (size_t)i * D + t));
        if (lane == 0) {
            atomicAdd(counts + lab, 1);
            local += (double)__uint_as_float((u32)(kx >> 32));
        }
    }
    if (lane == 0 && local != 0.0) atomicAdd(inertia, local);
}

// centres := sums / counts (an empty cluster keeps its centre); shift2 = sum of squared moves.
__global__ void kmeans_finish_kernel(const float *__restrict__ sums, const int *__restrict__ counts, long long k, int d,
                                     float *centers, double *shift2)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double mv = 0.0;
    if (i < k * d) {
        const int cnt = counts[i / d];
        const float old = centers[i];
        const float nw = cnt > 0 ? sums[i] / (float)cnt : old;
        centers[i] = nw;
        mv = (double)(nw - old) * (double)(nw - old);
    }
#pragma unroll
    for (int m = 16; m; m >>= 1) mv += __shfl_xor_sync(FULL_MASK, mv, m);
    if ((threadIdx.x & 31) == 0 && mv != 0.0) atomicAdd(shift2, mv);
}

}  // namespace

extern "C" int gp_kmeans_assign(const float *d_emb, const float *d_centers, int64_t num_nodes, int64_t num_centers,
                                int64_t dim, uint64_t *d_best, gp_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(num_nodes >= 0 && num_centers > 0, GP_ERR_INVALID, "gp_kmeans_assign: bad sizes");
    GP_REQUIRE(dim == 64 || dim == 128, GP_ERR_UNSUPPORTED, "gp_kmeans_assign: embedding dimension must be 64 or 128");
    if (num_nodes == 0) return GP_OK;
    GP_REQUIRE(d_emb != nullptr && d_centers != nullptr && d_best != nullptr, GP_ERR_INVALID, "gp_kmeans_assign: NULL argument");
    GP_LAUNCH(fill_u64_kernel, (unsigned)gp_ceil_div(num_nodes, 256), 256, 0, stream, (unsigned long long *)d_best,
              num_nodes, ~0ull);
    CdistParams p;
    memset(&p, 0, sizeof(p));
    p.emb = d_emb;
    p.anc = d_centers;
    p.n = num_nodes;
    p.k = num_centers;
    p.d = dim;
    p.mode = GP_CDIST_EUCLIDEAN;
    p.best = reinterpret_cast<unsigned long long *>(d_best);
    const int64_t col_tiles = gp_ceil_div(num_centers, CD_MAX_N);
    const int64_t ktile = num_centers < CD_MAX_N ? num_centers : CD_MAX_N;
    const int n_pad = (int)(gp_ceil_div(ktile, 32) * 32);
    const size_t smem = cdist_smem_bytes((int)dim, n_pad);
    const int64_t row_tiles = gp_ceil_div(num_nodes, CD_TILE_M);
    int grid_x = gp_sm_count();
    if (row_tiles < grid_x) grid_x = (int)row_tiles;
    const dim3 grid(grid_x, (unsigned)col_tiles);
    gp_count_launch();
    if (dim == 128) {
        GP_CUDA_CHECK(cudaFuncSetAttribute(cdist_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cdist_kernel<128><<<grid, CD_THREADS2, smem, stream>>>(p, n_pad, 0, 0, 0);
    } else {
        GP_CUDA_CHECK(cudaFuncSetAttribute(cdist_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cdist_kernel<64><<<grid, CD_THREADS2, smem, stream>>>(p, n_pad, 0, 0, 0);
    }
    GP_CUDA_CHECK(cudaGetLastError());
    return GP_OK;
}

extern "C" int gp_kmeans_update(const float *d_emb, const uint64_t *d_best, int64_t num_nodes, int64_t num_centers,
                                int64_t dim, float *d_centers, float *d_sums, int32_t *d_counts, double *d_shift2,
                                double *d_inertia, gp_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(dim == 64 || dim == 128, GP_ERR_UNSUPPORTED, "gp_kmeans_update: embedding dimension must be 64 or 128");
    GP_REQUIRE(d_emb && d_best && d_centers && d_sums && d_counts && d_shift2 && d_inertia, GP_ERR_INVALID,
               "gp_kmeans_update: NULL argument");
    GP_CUDA_CHECK(cudaMemsetAsync(d_sums, 0, sizeof(float) * (size_t)(num_centers * dim), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, sizeof(int) * (size_t)num_centers, stream));
    GP_CUDA_CHECK(cudaMemsetAsync(d_shift2, 0, sizeof(double), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(d_inertia, 0, sizeof(double), stream));
    const int blocks = (int)std::min<int64_t>(gp_ceil_div(num_nodes > 0 ? num_nodes : 1, 8), (int64_t)gp_sm_count() * 16);
    if (dim == 128)
        GP_LAUNCH(kmeans_accumulate_kernel<128>, blocks, 256, 0, stream, d_emb, num_nodes, (const unsigned long long *)d_best,
                  d_sums, d_counts, d_inertia);
    else
        GP_LAUNCH(kmeans_accumulate_kernel<64>, blocks, 256, 0, stream, d_emb, num_nodes, (const unsigned long long *)d_best,
                  d_sums, d_counts, d_inertia);
    GP_LAUNCH(kmeans_finish_kernel, (unsigned)gp_ceil_div(num_centers * dim, 256), 256, 0, stream, d_sums, d_counts,
              num_centers, (int)dim, d_centers, d_shift2);
    GP_CUDA_CHECK(cudaGetLastError());
    return GP_OK;
}

extern "C" int gp_kmeans_plusplus(const float *d_emb, int64_t num_nodes, int64_t num_centers, int64_t dim,
                                  int64_t first_index, uint64_t seed, float *d_centers, float *d_mind2,
                                  int64_t *d_chosen, gp_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(dim == 64 || dim == 128, GP_ERR_UNSUPPORTED, "gp_kmeans_plusplus: embedding dimension must be 64 or 128");
    GP_REQUIRE(num_nodes > 0 && num_centers > 0 && first_index >= 0 && first_index < num_nodes, GP_ERR_INVALID,
               "gp_kmeans_plusplus: bad sizes");
    GP_REQUIRE(d_emb && d_centers && d_mind2 && d_chosen, GP_ERR_INVALID, "gp_kmeans_plusplus: NULL argument");
    // greedy k-means++: 2 + log(K) candidates per step (scikit-learn's n_local_trials), the one that lowers
    // the potential most wins
    int trials = 2 + (int)log((double)num_centers);
    if (trials > KPP_MAX_TRIALS) trials = KPP_MAX_TRIALS;
    if (trials < 1) trials = 1;
    unsigned long long *key = nullptr;  // [KPP_MAX_TRIALS] race winners, then [KPP_MAX_TRIALS] doubles (potentials)
    GP_CUDA_CHECK(cudaMalloc((void **)&key, 2 * KPP_MAX_TRIALS * sizeof(unsigned long long)));
    double *pot = reinterpret_cast<double *>(key + KPP_MAX_TRIALS);
    long long first = first_index;
    cudaError_t e = cudaMemcpyAsync(d_chosen, &first, sizeof(long long), cudaMemcpyHostToDevice, stream);
    const int blocks = (int)std::min<int64_t>(gp_ceil_div(num_nodes, 8), (int64_t)gp_sm_count() * 16);
    for (int64_t s = 0; e == cudaSuccess && s < num_centers; ++s) {
        GP_LAUNCH(fill_u64_kernel, 1, 32, 0, stream, key, (long long)KPP_MAX_TRIALS, ~0ull);
        e = cudaMemsetAsync(pot, 0, KPP_MAX_TRIALS * sizeof(double), stream);
        if (e != cudaSuccess) break;
        if (dim == 128)
            GP_LAUNCH(kpp_step_kernel<128>, blocks, 256, 0, stream, d_emb, num_nodes, (const long long *)d_chosen, (int)s,
                      (int)num_centers, trials, (u64)seed, d_mind2, d_centers, key);
        else
            GP_LAUNCH(kpp_step_kernel<64>, blocks, 256, 0, stream, d_emb, num_nodes, (const long long *)d_chosen, (int)s,
                      (int)num_centers, trials, (u64)seed, d_mind2, d_centers, key);
        if (s + 1 < num_centers) {
            if (dim == 128)
                GP_LAUNCH(kpp_potential_kernel<128>, blocks, 256, 0, stream, d_emb, num_nodes, key, trials, d_mind2, pot);
            else
                GP_LAUNCH(kpp_potential_kernel<64>, blocks, 256, 0, stream, d_emb, num_nodes, key, trials, d_mind2, pot);
            GP_LAUNCH(kpp_pick_kernel, 1, 1, 0, stream, key, pot, trials, (long long *)d_chosen, (int)s);
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // `first` and `key` live until the stream is done
    cudaFree(key);
    GP_CUDA_CHECK(e);
    return GP_OK;
}
