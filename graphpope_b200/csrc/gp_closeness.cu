// gp_closeness.cu — closeness-centrality scores for the 'closeness_centrality' anchor sampler
// (reference utils.py:50-54: nx.closeness_centrality(G), then the stable top-k of utils.py:53-54).
//
// networkx (closeness_centrality, wf_improved=True, incoming distance on a DiGraph): for node a,
//   sp     = hop counts hops(u -> a) of every node u that can reach a (a itself included, 0),
//   c(a)   = (len(sp) - 1) / sum(sp)  *  (len(sp) - 1) / (N - 1)        (0 if sum(sp) == 0 or N == 1).
// hops(u -> a) for all u is exactly one column of the anchor-distance matrix, so the scores of ALL
// nodes are the MS-BFS of gp_msbfs.cu run with every node as an anchor, 4096 anchors per pass, followed
// by two column sums per anchor (reached count, hop total).  The sums are taken straight from the
// bit-sliced result block: per (row, lane word) the hop index is re-assembled as bit planes (the same
// OR-over-levels the decode epilogue does) and added into bit-sliced per-lane counters with a
// ripple-carry adder made of bitwise ops, 64 anchors per instruction; nothing is ever expanded to a
// per-(node, anchor) integer.  All arithmetic on the scores is float64 in networkx's operation order,
// so they are bit-equal to the reference's.
#include "gp_msbfs.cuh"

namespace {

constexpr int CC_ROWS = 64;        // words per lane and chunk (bit-sliced counters must hold CC_ROWS * max hop)
constexpr int CC_THREADS = 256;

__global__ void iota_i64_kernel(long long *a, long long first, long long k)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) a[i] = first + i;
}

// P bit planes: c += h (h given as H planes), all 64 lanes of a word at once.
template <int P, int H>
__device__ __forceinline__ void bitsliced_add(u64 (&c)[P], const u64 (&h)[H])
{
    u64 carry = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const u64 a = c[p], b = p < H ? h[p] : 0ull;
        const u64 axb = a ^ b;
        c[p] = axb ^ carry;
        carry = (a & b) | (carry & axb);
    }
}

// sums[0][lane] += rows that reached the lane's anchor; sums[1][lane] += total hops to it.
// One warp = CC_ROWS * 32 consecutive lane words of one batch (coalesced; a lane keeps its word slot).
template <bool DEEP>
__global__ void __launch_bounds__(CC_THREADS)
colsum_kernel(const u64 *__restrict__ result, long long plane_stride, const int *__restrict__ status, long long n,
              int wb, int batches, unsigned long long *sums, long long lane_cap)
{
    constexpr int P = DEEP ? 24 : 12;  // 64 rows * 65534 hops < 2^22; 64 * 15 < 2^10
    constexpr int H = DEEP ? 16 : 4;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long words_per_batch = n * wb;
    const long long chunks_per_batch = (words_per_batch + CC_ROWS * 32 - 1) / (CC_ROWS * 32);
    if (warp >= chunks_per_batch * batches) return;
    const int b = (int)(warp / chunks_per_batch);
    const long long base = (warp - (long long)b * chunks_per_batch) * (CC_ROWS * 32);
    const int max_level = status[GP_BFS_ST_MAX_LEVEL];
    const int levels = max_level < GP_BFS_LEVEL_ARRAYS ? max_level : GP_BFS_LEVEL_ARRAYS;
    const u64 *blk = result + (size_t)b * words_per_batch;

    u64 tot[P], cnt[8];
#pragma unroll
    for (int p = 0; p < P; ++p) tot[p] = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) cnt[p] = 0;
    for (int it = 0; it < CC_ROWS; ++it) {
        const long long i = base + (long long)it * 32 + lane;
        if (i >= words_per_batch) break;
        const u64 reach[1] = {blk[i]};
        u64 h[H];
#pragma unroll
        for (int q = 0; q < H; ++q) h[q] = 0;
        for (int l = 1; l <= levels; ++l) {  // hop index of the shallow lanes, bit-sliced
            const u64 x = blk[(size_t)l * plane_stride + i];
            if (l & 1) h[0] |= x;
            if (l & 2) h[1] |= x;
            if (l & 4) h[2] |= x;
            if (l & 8) h[3] |= x;
        }
        if (DEEP) {
#pragma unroll
            for (int q = 0; q < GP_BFS_PLANES; ++q)
                h[q] |= blk[(size_t)(1 + GP_BFS_LEVEL_ARRAYS + q) * plane_stride + i];
        }
        bitsliced_add<P, H>(tot, h);
        bitsliced_add<8, 1>(cnt, reach);
    }
    // expand the counters of the 64 lanes of this word slot, fold the lanes that share a slot, publish
    const int w = lane % wb;
    for (int j = 0; j < 64; ++j) {
        u32 t = 0, c = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) t |= (u32)((tot[p] >> j) & 1ull) << p;
#pragma unroll
        for (int p = 0; p < 8; ++p) c |= (u32)((cnt[p] >> j) & 1ull) << p;
        for (int m = wb; m < 32; m <<= 1) {
            t += __shfl_xor_sync(FULL_MASK, t, m);
            c += __shfl_xor_sync(FULL_MASK, c, m);
        }
        if (lane < wb) {
            const long long col = ((long long)b * wb + w) * 64 + j;
            if (c) atomicAdd(sums + col, (unsigned long long)c);
            if (t) atomicAdd(sums + lane_cap + col, (unsigned long long)t);
        }
    }
}

// networkx operation order, float64: cc = (r - 1) / totsp; s = (r - 1) / (N - 1); cc *= s.
__global__ void closeness_finish_kernel(const unsigned long long *__restrict__ sums, long long lane_cap, long long k,
                                        long long n, double *__restrict__ score)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    const double r = (double)sums[j], totsp = (double)sums[lane_cap + j];
    double cc = 0.0;
    if (totsp > 0.0 && n > 1) {
        cc = __ddiv_rn(__dsub_rn(r, 1.0), totsp);
        const double s = __ddiv_rn(__dsub_rn(r, 1.0), (double)(n - 1));
        cc = __dmul_rn(cc, s);
    }
    score[j] = cc;
}

}  // namespace

extern "C" int gp_closeness(const gp_csr_t *csr_, double *d_score, gp_stream_t stream_)
{
    gp_csr *csr = const_cast<gp_csr *>(csr_);
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(csr != nullptr && d_score != nullptr, GP_ERR_INVALID, "gp_closeness: NULL argument");
    GP_REQUIRE(csr->built, GP_ERR_INVALID, "gp_closeness: the CSR has not been built");
    const long long n = csr->num_nodes;
    if (n == 0) return GP_OK;
    const long long chunk = n < 4096 ? n : 4096;
    gp_msbfs_t *h = nullptr;
    GP_TRY(gp_msbfs_create(csr, chunk, &h));
    const long long lane_cap = (chunk + 255) / 256 * 256;
    long long *d_anchors = nullptr;
    unsigned long long *d_sums = nullptr;
    int rc = GP_OK;
    if (cudaMalloc((void **)&d_anchors, sizeof(long long) * (size_t)chunk) != cudaSuccess ||
        cudaMalloc((void **)&d_sums, sizeof(unsigned long long) * 2 * (size_t)lane_cap) != cudaSuccess) {
        gp_set_error("gp_closeness: cudaMalloc failed");
        rc = GP_ERR_OOM;
    }
    for (long long c0 = 0; rc == GP_OK && c0 < n; c0 += chunk) {
        const long long k = n - c0 < chunk ? n - c0 : chunk;
        GP_LAUNCH(iota_i64_kernel, (unsigned)gp_ceil_div(k, 256), 256, 0, stream, d_anchors, c0, k);
        rc = gp_msbfs_run(h, (const int64_t *)d_anchors, k, stream);
        if (rc != GP_OK) break;
        if (cudaMemsetAsync(d_sums, 0, sizeof(unsigned long long) * 2 * (size_t)lane_cap, stream) != cudaSuccess) {
            gp_set_error("gp_closeness: cudaMemsetAsync failed");
            rc = GP_ERR_CUDA;
            break;
        }
        // the hop-depth decides the adder width; it is only known on the device, so the shallow kernel
        // bails out (nothing published) when the run went deeper than 15 hops and the deep one runs then
        gp_msbfs_stats_t st;
        rc = gp_msbfs_stats(h, &st, stream);  // syncs; also surfaces index errors
        if (rc != GP_OK) break;
        const long long words_per_batch = n * h->wb;
        const long long chunks = gp_ceil_div(words_per_batch, CC_ROWS * 32) * h->batches;
        const unsigned blocks = (unsigned)gp_ceil_div(chunks * 32, CC_THREADS);
        const long long plane_stride = (long long)h->wb * h->batches * n;
        if (st.max_level > GP_BFS_LEVEL_ARRAYS)
            GP_LAUNCH(colsum_kernel<true>, blocks, CC_THREADS, 0, stream, h->seen, plane_stride, h->status, n, h->wb,
                      h->batches, d_sums, lane_cap);
        else
            GP_LAUNCH(colsum_kernel<false>, blocks, CC_THREADS, 0, stream, h->seen, plane_stride, h->status, n, h->wb,
                      h->batches, d_sums, lane_cap);
        GP_LAUNCH(closeness_finish_kernel, (unsigned)gp_ceil_div(k, 256), 256, 0, stream, d_sums, lane_cap, k, n,
                  d_score + c0);
        if (cudaGetLastError() != cudaSuccess) {
            gp_set_error("gp_closeness: kernel launch failed");
            rc = GP_ERR_CUDA;
        }
    }
    if (rc == GP_OK && cudaStreamSynchronize(stream) != cudaSuccess) {
        gp_set_error("gp_closeness: stream synchronisation failed");
        rc = GP_ERR_CUDA;
    }
    cudaFree(d_anchors);
    cudaFree(d_sums);
    gp_msbfs_free(h);
    return rc;
}

// ================================================================= clustering coefficient (utils.py:56-60)
// nx.clustering(G) on the DiGraph (Fagiolo's directed clustering): for node i with predecessors P and
// successors S (self-loops dropped),
//   t  = sum over j in chain(P, S) of |P & Pj| + |P & Sj| + |S & Pj| + |S & Sj|
//   c  = 0 if t == 0 else t / (2 * (dt * (dt - 1) - 2 * db)),   dt = |P| + |S|, db = |P & S|.
// One CTA per node: P and S become two node bitmaps in shared memory, then every (j, k in Pj or Sj) pair
// is two bit tests; all counts are integers and the one division is float64, so the scores are bit-equal
// to networkx's.  Graphs whose two bitmaps do not fit in shared memory are not supported (the host layer
// then keeps the reference's networkx call).
namespace {

constexpr int CL_THREADS = 256;

__global__ void __launch_bounds__(CL_THREADS)
clustering_kernel(const int *__restrict__ row_start, const int *__restrict__ deg, const int *__restrict__ col,
                  const int *__restrict__ rowptr_in, const int *__restrict__ col_in, long long n, int words,
                  double *__restrict__ score)
{
    extern __shared__ u32 s_bits[];  // [2][words]: membership in P, in S
    __shared__ unsigned long long s_t;
    __shared__ int s_db, s_dt;
    u32 *bp = s_bits, *bs = s_bits + words;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int w = tid; w < 2 * words; w += CL_THREADS) s_bits[w] = 0;
    __syncthreads();
    for (long long i = blockIdx.x; i < n; i += gridDim.x) {
        const int s0 = row_start[i], sd = deg[i];            // successors: col[s0 .. s0 + sd)
        const int p0 = rowptr_in[i], pd = rowptr_in[i + 1] - p0;  // predecessors: col_in[p0 .. p0 + pd)
        if (tid == 0) {
            s_t = 0;
            s_db = 0;
            s_dt = 0;
        }
        for (int e = tid; e < pd; e += CL_THREADS) {
            const int v = col_in[p0 + e];
            if (v != i) atomicOr(&bp[v >> 5], 1u << (v & 31));
        }
        for (int e = tid; e < sd; e += CL_THREADS) {
            const int v = col[s0 + e];
            if (v != i) atomicOr(&bs[v >> 5], 1u << (v & 31));
        }
        __syncthreads();
        // dt, db
        int dt = 0, db = 0;
        for (int e = tid; e < pd; e += CL_THREADS) dt += col_in[p0 + e] != i;
        for (int e = tid; e < sd; e += CL_THREADS) {
            const int v = col[s0 + e];
            if (v != i) {
                ++dt;
                db += (bp[v >> 5] >> (v & 31)) & 1u;
            }
        }
        dt = __reduce_add_sync(FULL_MASK, dt);
        db = __reduce_add_sync(FULL_MASK, db);
        if (lane == 0) {
            atomicAdd(&s_dt, dt);
            atomicAdd(&s_db, db);
        }
        // triangles: one warp per j of chain(P, S), lanes over the neighbours k of j
        unsigned long long t = 0;
        for (int e = warp; e < pd + sd; e += CL_THREADS / 32) {
            const int j = e < pd ? col_in[p0 + e] : col[s0 + (e - pd)];
            if (j == i) continue;
            const int js0 = row_start[j], jsd = deg[j];
            const int jp0 = rowptr_in[j], jpd = rowptr_in[j + 1] - jp0;
            for (int q = lane; q < jpd + jsd; q += 32) {
                const int k = q < jpd ? col_in[jp0 + q] : col[js0 + (q - jpd)];
                if (k != j) t += ((bp[k >> 5] >> (k & 31)) & 1u) + ((bs[k >> 5] >> (k & 31)) & 1u);
            }
        }
        for (int m = 16; m; m >>= 1) t += shfl_xor_u64(t, m);
        if (lane == 0 && t) atomicAdd(&s_t, t);
        __syncthreads();
        if (tid == 0) {
            const double tt = (double)s_t;
            const long long dtl = s_dt, dbl = s_db;
            score[i] = s_t == 0 ? 0.0 : __ddiv_rn(tt, (double)((dtl * (dtl - 1) - 2 * dbl) * 2));
        }
        // clear exactly the bits that were set
        for (int e = tid; e < pd; e += CL_THREADS) bp[col_in[p0 + e] >> 5] = 0;
        for (int e = tid; e < sd; e += CL_THREADS) bs[col[s0 + e] >> 5] = 0;
        __syncthreads();
    }
}

}  // namespace

extern "C" int gp_clustering(const gp_csr_t *csr_, double *d_score, gp_stream_t stream_)
{
    gp_csr *csr = const_cast<gp_csr *>(csr_);
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(csr != nullptr && d_score != nullptr, GP_ERR_INVALID, "gp_clustering: NULL argument");
    GP_REQUIRE(csr->built, GP_ERR_INVALID, "gp_clustering: the CSR has not been built");
    const long long n = csr->num_nodes;
    if (n == 0) return GP_OK;
    const int words = (int)((n + 31) / 32);
    const size_t smem = 2 * (size_t)words * sizeof(u32);
    GP_REQUIRE(smem <= 200 * 1024, GP_ERR_UNSUPPORTED,
               "gp_clustering: two %lld-node bitmaps do not fit in shared memory", n);
    GP_TRY(gp_csr_ensure_in(csr, stream));
    GP_CUDA_CHECK(cudaFuncSetAttribute(clustering_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)(200 * 1024 / (smem > 0 ? smem : 1));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)gp_sm_count() * per_sm;
    if (blocks > n) blocks = n;
    GP_LAUNCH(clustering_kernel, (unsigned)blocks, CL_THREADS, smem, stream, csr->row_start, csr->deg, csr->col,
              csr->rowptr_in, csr->col_in, n, words, d_score);
    GP_CUDA_CHECK(cudaGetLastError());
    GP_CUDA_CHECK(cudaStreamSynchronize(stream));
    return GP_OK;
}
