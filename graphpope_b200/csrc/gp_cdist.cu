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
// Tiling: one CTA = 128 nodes x (<= 256 anchors); the anchor operand (hi and lo, K-major,
// 128-byte swizzle) stays resident in shared memory for the CTA's lifetime, node tiles are staged
// tile by tile; accumulators are 128 TMEM lanes x 256 columns.  The kernel is HBM bound by
// construction (AI ~ 43 flop/B at D = 128): 4*N*D bytes in, 4*N*K bytes out.
#include "gp_internal.h"

#include <cuda_bf16.h>

#include <algorithm>

namespace {

constexpr int CD_TILE_M = 128;
constexpr int CD_THREADS = 128;
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
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// Orderable-int encoding so float min/max can use integer atomics for any sign.
__device__ __forceinline__ int f2ord(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// Stage `rows` fp32 rows of length D (row-major, leading dim D) as bf16 hi / lo operands in the
// canonical K-major SWIZZLE_128B layout: [D/64 chunks][rows_pad][128 B], 16-byte column c of row r
// stored at column c ^ (r & 7).  Rows >= valid are zero.  Also writes the exact fp32 row norms.
template <int D>
__device__ __forceinline__ void stage_operand(const float *__restrict__ src, long long first_row, long long valid,
                                              int rows_pad, unsigned char *s_hi, unsigned char *s_lo,
                                              float *s_norm)
{
    constexpr int GROUPS = D / 8;  // 8-element (16-byte bf16) groups per row
    const int total = rows_pad * GROUPS;
    for (int g = threadIdx.x; g < total; g += CD_THREADS) {
        const int row = g / GROUPS, kg = g % GROUPS;
        float x[8];
        if ((long long)row < valid) {
            const float4 *p = reinterpret_cast<const float4 *>(src + (size_t)(first_row + row) * D + kg * 8);
            const float4 a = __ldg(p), b = __ldg(p + 1);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
            x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 0.0f;
        }
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
        const int chunk = kg / 8, c = kg % 8;
        const size_t off = (size_t)chunk * rows_pad * 128 + (size_t)row * 128 + (size_t)((c ^ (row & 7)) * 16);
        *reinterpret_cast<uint4 *>(s_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(s_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        // row norm: the GROUPS threads of a row are consecutive lanes (GROUPS = 8 or 16 divides 32)
#pragma unroll
        for (int m = GROUPS / 2; m; m >>= 1) ss += __shfl_xor_sync(FULL_MASK, ss, m);
        if (kg == 0) s_norm[row] = ss;
    }
}

template <int D>
__global__ void __launch_bounds__(CD_THREADS, 1) cdist_kernel(CdistParams p, int n_pad)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: B hi | B lo | A hi | A lo (each 1024-aligned), then norms + barrier + tmem address
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const size_t b_bytes = (size_t)(D / CD_KCHUNK) * n_pad * 128;
    const size_t a_bytes = (size_t)(D / CD_KCHUNK) * CD_TILE_M * 128;
    unsigned char *sB_hi = smem, *sB_lo = sB_hi + b_bytes;
    unsigned char *sA_hi = sB_lo + b_bytes, *sA_lo = sA_hi + a_bytes;
    float *s_an = reinterpret_cast<float *>(sA_lo + a_bytes);
    float *s_xn = s_an + CD_MAX_N;
    u64 *s_bar = reinterpret_cast<u64 *>(s_xn + CD_TILE_M);
    u32 *s_tmem = reinterpret_cast<u32 *>(s_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long col_tile = blockIdx.y;
    const long long k0 = col_tile * CD_MAX_N;
    const long long kvalid = min((long long)CD_MAX_N, p.k - k0);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(s_bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // anchors of this column tile: resident for the CTA's lifetime
    stage_operand<D>(p.anc, k0, kvalid, n_pad, sB_hi, sB_lo, s_an);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_base = *s_tmem;
    const u32 idesc = make_idesc(CD_TILE_M, n_pad);
    const u32 bar = smem_u32(s_bar);
    u32 parity = 0;

    const long long row_tiles = (p.n + CD_TILE_M - 1) / CD_TILE_M;
    for (long long tile = blockIdx.x; tile < row_tiles; tile += gridDim.x) {
        const long long r0 = tile * CD_TILE_M;
        const long long rvalid = min((long long)CD_TILE_M, p.n - r0);
        stage_operand<D>(p.emb, r0, rvalid, CD_TILE_M, sA_hi, sA_lo, s_xn);
        // generic-proxy smem writes -> visible to the tensor core's async proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
                        umma_bf16(tmem_base, da, db, idesc, acc);
                        acc = 1;
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                         : "memory");
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // epilogue: thread = one node row (TMEM lane), 16 anchor columns per tcgen05.ld
        const int rloc = warp * 32 + lane;
        const long long row = r0 + rloc;
        const float xx = s_xn[rloc];
        const float inv_xn = xx > 0.0f ? rsqrtf(xx) : 0.0f;
        for (int c0 = 0; c0 < n_pad; c0 += 16) {
            u32 v[16];
            const u32 taddr = tmem_base + ((u32)(warp * 32) << 16) + (u32)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                "%14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < p.n) {
                float res[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = c0 + j;
                    const float dot = __uint_as_float(v[j]);
                    const float aa = s_an[c];
                    float r;
                    if (p.mode == GP_CDIST_EUCLIDEAN) {
                        float d2 = xx - 2.0f * dot + aa;
                        if (d2 < 1.0e-3f * (xx + aa) && c < kvalid) {
                            // cancellation zone: recompute from differences (exact 0 for identical rows)
                            const float *xr = p.emb + (size_t)row * D, *ar = p.anc + (size_t)(k0 + c) * D;
                            float s = 0.0f;
                            for (int t = 0; t < D; ++t) {
                                const float df = __ldg(xr + t) - __ldg(ar + t);
                                s = fmaf(df, df, s);
                            }
                            d2 = s;
                        }
                        r = sqrtf(fmaxf(d2, 0.0f));
                    } else {
                        const float inv_an = aa > 0.0f ? rsqrtf(aa) : 0.0f;
                        const float sim = dot * inv_xn * inv_an;
                        r = p.mode == GP_CDIST_COSINE_SIMILARITY ? sim : fminf(fmaxf(1.0f - sim, 0.0f), 2.0f);
                    }
                    res[j] = r;
                }
                float *orow = p.out + (size_t)row * p.ld_out + p.col_offset + k0 + c0;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < kvalid) orow[j] = res[j];
            }
        }
        // all warps have drained TMEM and sA before the next tile overwrites them
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// Column min / max of out[:, col_offset : col_offset + k] (MinMaxScaler.fit, utils.py:175).
__global__ void __launch_bounds__(256) col_minmax_kernel(const float *__restrict__ out, long long n, long long k,
                                                         long long ld, long long col_offset, int *cmin, int *cmax)
{
    const long long rows_per_block = (n + gridDim.y - 1) / gridDim.y;
    const long long r_begin = (long long)blockIdx.y * rows_per_block;
    const long long r_end = min(n, r_begin + rows_per_block);
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k || r_begin >= r_end) return;
    float mn = INFINITY, mx = -INFINITY;
    for (long long r = r_begin; r < r_end; ++r) {
        const float v = out[(size_t)r * ld + col_offset + c];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    atomicMin(cmin + c, f2ord(mn));
    atomicMax(cmax + c, f2ord(mx));
}

__global__ void minmax_init_kernel(int *cmin, int *cmax, long long k)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < k) {
        cmin[c] = f2ord(INFINITY);
        cmax[c] = f2ord(-INFINITY);
    }
}

// MinMaxScaler.transform (utils.py:176): X * scale + min_, scale = 1/range (range < 10 eps -> 1).
__global__ void __launch_bounds__(256) col_scale_kernel(float *__restrict__ out, long long n, long long k,
                                                        long long ld, long long col_offset,
                                                        const int *__restrict__ cmin, const int *__restrict__ cmax)
{
    const long long total = n * k;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / k, c = i - r * k;
        const float dmin = ord2f(cmin[c]), dmax = ord2f(cmax[c]);
        float range = dmax - dmin;
        if (range < 10.0f * 1.1920929e-07f) range = 1.0f;  // sklearn _handle_zeros_in_scale
        const float scale = __fdiv_rn(1.0f, range);
        const float mn = __fsub_rn(0.0f, __fmul_rn(dmin, scale));
        float *q = out + (size_t)r * ld + col_offset + c;
        *q = __fadd_rn(__fmul_rn(*q, scale), mn);
    }
}

size_t cdist_smem_bytes(int d, int n_pad)
{
    const size_t b = (size_t)(d / CD_KCHUNK) * n_pad * 128, a = (size_t)(d / CD_KCHUNK) * CD_TILE_M * 128;
    return 1024 + 2 * b + 2 * a + sizeof(float) * (CD_MAX_N + CD_TILE_M) + 64;
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
    GP_REQUIRE(dim == 64 || dim == 128, GP_ERR_UNSUPPORTED,
               "gp_cdist_minmax: embedding dimension %lld not supported by this build (64 or 128; the reference "
               "uses 128, generate_node2vec_embedding.py:23)", (long long)dim);
    GP_REQUIRE((reinterpret_cast<uintptr_t>(d_emb) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d_anchor_emb) & 15u) == 0,
               GP_ERR_INVALID, "gp_cdist_minmax: embeddings must be 16-byte aligned");
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
    const int64_t col_tiles = gp_ceil_div(num_anchors, CD_MAX_N);
    const int64_t ktile = num_anchors < CD_MAX_N ? num_anchors : CD_MAX_N;
    const int n_pad = (int)(gp_ceil_div(ktile, 16) * 16);  // UMMA N: multiple of 16 at M = 128
    const size_t smem = cdist_smem_bytes((int)dim, n_pad);
    const int64_t row_tiles = gp_ceil_div(num_nodes, CD_TILE_M);
    int grid_x = gp_sm_count();
    if (row_tiles < grid_x) grid_x = (int)row_tiles;
    gp_count_launch();
    if (dim == 128) {
        GP_CUDA_CHECK(cudaFuncSetAttribute(cdist_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cdist_kernel<128><<<dim3(grid_x, (unsigned)col_tiles), CD_THREADS, smem, stream>>>(p, n_pad);
    } else {
        GP_CUDA_CHECK(cudaFuncSetAttribute(cdist_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cdist_kernel<64><<<dim3(grid_x, (unsigned)col_tiles), CD_THREADS, smem, stream>>>(p, n_pad);
    }
    GP_CUDA_CHECK(cudaGetLastError());
    if (apply_minmax) {
        int *mm = nullptr;
        GP_CUDA_CHECK(cudaMallocAsync((void **)&mm, sizeof(int) * 2 * (size_t)num_anchors, stream));
        int *cmin = mm, *cmax = mm + num_anchors;
        GP_LAUNCH(minmax_init_kernel, (unsigned)gp_ceil_div(num_anchors, 256), 256, 0, stream, cmin, cmax, num_anchors);
        const int ysplit = (int)std::min<int64_t>(std::max<int64_t>(1, num_nodes / 256), (int64_t)gp_sm_count() * 2);
        GP_LAUNCH(col_minmax_kernel, dim3((unsigned)gp_ceil_div(num_anchors, 256), ysplit), 256, 0, stream, d_out,
                  num_nodes, num_anchors, ld_out, col_offset, cmin, cmax);
        int64_t blocks = gp_ceil_div(num_nodes * num_anchors, 256);
        if (blocks > (int64_t)gp_sm_count() * 16) blocks = (int64_t)gp_sm_count() * 16;
        GP_LAUNCH(col_scale_kernel, (unsigned)blocks, 256, 0, stream, d_out, num_nodes, num_anchors, ld_out,
                  col_offset, cmin, cmax);
        GP_CUDA_CHECK(cudaFreeAsync(mm, stream));
        GP_CUDA_CHECK(cudaGetLastError());
    }
    return GP_OK;
}
