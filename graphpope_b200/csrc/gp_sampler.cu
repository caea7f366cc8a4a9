// gp_sampler.cu — centrality-biased anchor sampling on the device (reference utils.py:26-30,
// 38-42): degree via row-pointer differences of the out- and in-edge CSR (the segmented
// reduction is already materialised there), PageRank via a float64 CSR pull over in-edges, and
// the stable "ascending sort, keep the last K" selection.
//
// PageRank follows networkx _pagerank_scipy operation by operation so the scores are
// bit-identical to the reference's: y[v] accumulates (1/outdeg[u]) * x[u] over the
// in-neighbours u of v in ascending u (scipy csc_matvec order), multiply then add with no
// FMA contraction; the dangling mass is a left-to-right sum (Python's sum()); then
//     x = alpha * (y + dsum * p) + (1 - alpha) * p,   stop when sum|x - xlast| < N * tol.
#include "gp_internal.h"
#include "gp_sort.cuh"

#include <vector>

namespace {

__global__ void degree_kernel(const int *__restrict__ deg_out, const int *__restrict__ rp_in, long long n,
                              int *__restrict__ deg)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += stride)
        deg[u] = deg_out[u] + (rp_in[u + 1] - rp_in[u]);
}

__global__ void pr_init_kernel(const int *__restrict__ deg_out, long long n, double *__restrict__ x,
                               double *__restrict__ inv_out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double x0 = 1.0 / (double)n;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += stride) {
        const int d = deg_out[u];
        inv_out[u] = d > 0 ? 1.0 / (double)d : 0.0;
        x[u] = x0;
    }
}

// Ascending list of nodes without out-edges (one warp, ballot compaction keeps the order).
__global__ void dangling_list_kernel(const int *__restrict__ deg_out, long long n, int *__restrict__ list,
                                     int *count)
{
    const int lane = threadIdx.x;
    int base = 0;
    for (long long u0 = 0; u0 < n; u0 += 32) {
        const long long u = u0 + lane;
        const bool d = u < n && deg_out[u] == 0;
        const u32 m = __ballot_sync(FULL_MASK, d);
        if (d) list[base + __popc(m & ((1u << lane) - 1u))] = (int)u;
        base += __popc(m);
    }
    if (lane == 0) *count = base;
}

// contrib[u] = inv_out[u] * x[u]; one extra warp folds the dangling mass left to right.
__global__ void pr_contrib_kernel(const double *__restrict__ x, const double *__restrict__ inv_out, long long n,
                                  double *__restrict__ contrib, const int *__restrict__ dangling,
                                  const int *__restrict__ n_dangling, double *dsum)
{
    if (blockIdx.x == gridDim.x - 1) {
        if (threadIdx.x >= 32) return;
        const int lane = threadIdx.x, m = *n_dangling;
        double s = 0.0;
        for (int i0 = 0; i0 < m; i0 += 32) {
            const double v = (i0 + lane < m) ? x[dangling[i0 + lane]] : 0.0;
            const int cnt = min(32, m - i0);
            for (int t = 0; t < cnt; ++t) s = __dadd_rn(s, __shfl_sync(FULL_MASK, v, t));
        }
        if (lane == 0) *dsum = s;
        return;
    }
    const long long stride = (long long)(gridDim.x - 1) * blockDim.x;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += stride)
        contrib[u] = __dmul_rn(inv_out[u], x[u]);
}

// One thread per destination row, strictly sequential accumulation (exact scipy order).
__global__ void __launch_bounds__(256)
pr_pull_kernel(const int *__restrict__ rp_in, const int *__restrict__ col_in, const double *__restrict__ contrib,
               const double *__restrict__ x, long long n, double alpha, double one_minus_alpha,
               const double *__restrict__ dsum, double *__restrict__ xnew, double *__restrict__ partial)
{
    __shared__ double s_part[8];
    const double p = 1.0 / (double)n;
    const double dterm = __dmul_rn(*dsum, p);
    const double base = __dmul_rn(one_minus_alpha, p);
    double err = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const int s = rp_in[v], e = rp_in[v + 1];
        double y = 0.0;
        int j = s;
        for (; j + 4 <= e; j += 4) {
            const double c0 = contrib[col_in[j]], c1 = contrib[col_in[j + 1]];
            const double c2 = contrib[col_in[j + 2]], c3 = contrib[col_in[j + 3]];
            y = __dadd_rn(y, c0);
            y = __dadd_rn(y, c1);
            y = __dadd_rn(y, c2);
            y = __dadd_rn(y, c3);
        }
        for (; j < e; ++j) y = __dadd_rn(y, contrib[col_in[j]]);
        const double xv = __dadd_rn(__dmul_rn(alpha, __dadd_rn(y, dterm)), base);
        xnew[v] = xv;
        err += fabs(xv - x[v]);
    }
    for (int m = 16; m; m >>= 1) err += __shfl_xor_sync(FULL_MASK, err, m);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = err;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void pr_err_kernel(const double *__restrict__ partial, int nblocks, double *err)
{
    // fixed-order final reduction: deterministic run to run
    __shared__ double s[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) t += partial[i];
    s[threadIdx.x] = t;
    __syncthreads();
    for (int m = 128; m; m >>= 1) {
        if ((int)threadIdx.x < m) s[threadIdx.x] += s[threadIdx.x + m];
        __syncthreads();
    }
    if (threadIdx.x == 0) *err = s[0];
}

__device__ __forceinline__ u64 orderable_f64(double v)
{
    const u64 b = (u64)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void topk_keys_i32_kernel(const int *__restrict__ score, long long n, u64 *__restrict__ keys)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += stride)
        keys[u] = ((u64)((u32)score[u] ^ 0x80000000u) << 32) | (u64)(u32)u;
}

__global__ void topk_keys_f64_kernel(const double *__restrict__ score, long long n, u64 *__restrict__ keys,
                                     u32 *__restrict__ vals)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += stride) {
        keys[u] = orderable_f64(score[u]);
        vals[u] = (u32)u;
    }
}

__global__ void topk_emit_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ vals, long long n,
                                 long long take, long long *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < take; i += stride) {
        const long long src = n - take + i;
        out[i] = vals ? (long long)vals[src] : (long long)(u32)keys[src];
    }
}

int blocks_for(long long n)
{
    long long b = gp_ceil_div(n > 0 ? n : 1, 256);
    const long long cap = (long long)gp_sm_count() * 8;
    return (int)(b < cap ? b : cap);
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" int gp_degree(const gp_csr_t *csr_, int32_t *d_degree, gp_stream_t stream_)
{
    gp_csr *csr = const_cast<gp_csr *>(csr_);
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(csr != nullptr && d_degree != nullptr, GP_ERR_INVALID, "gp_degree: NULL argument");
    GP_REQUIRE(csr->built, GP_ERR_INVALID, "gp_degree: the CSR has not been built");
    GP_TRY(gp_csr_ensure_in(csr, stream));
    if (csr->num_nodes == 0) return GP_OK;
    GP_LAUNCH(degree_kernel, blocks_for(csr->num_nodes), 256, 0, stream, csr->deg, csr->rowptr_in,
                                                                  csr->num_nodes, d_degree);
    GP_CUDA_CHECK(cudaGetLastError());
    return GP_OK;
}

extern "C" int gp_pagerank(const gp_csr_t *csr_, double alpha, double tol, int32_t max_iter, double *d_x,
                           int32_t *iterations, gp_stream_t stream_)
{
    gp_csr *csr = const_cast<gp_csr *>(csr_);
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(csr != nullptr && d_x != nullptr, GP_ERR_INVALID, "gp_pagerank: NULL argument");
    GP_REQUIRE(csr->built, GP_ERR_INVALID, "gp_pagerank: the CSR has not been built");
    if (iterations) *iterations = 0;
    const long long n = csr->num_nodes;
    if (n == 0) return GP_OK;
    GP_TRY(gp_csr_ensure_in(csr, stream));
    const int nblocks = blocks_for(n);
    DevBuf b_inv, b_contrib, b_xa, b_dang, b_small, b_partial;
    GP_CUDA_CHECK(cudaMalloc(&b_inv.p, sizeof(double) * n));
    GP_CUDA_CHECK(cudaMalloc(&b_contrib.p, sizeof(double) * n));
    GP_CUDA_CHECK(cudaMalloc(&b_xa.p, sizeof(double) * n));
    GP_CUDA_CHECK(cudaMalloc(&b_dang.p, sizeof(int) * (n + 1)));
    GP_CUDA_CHECK(cudaMalloc(&b_small.p, 64));
    GP_CUDA_CHECK(cudaMalloc(&b_partial.p, sizeof(double) * nblocks));
    double *inv_out = (double *)b_inv.p, *contrib = (double *)b_contrib.p;
    double *xa = d_x, *xb = (double *)b_xa.p;
    int *dang = (int *)b_dang.p;
    double *dsum = (double *)b_small.p, *err = dsum + 1;
    int *n_dang = (int *)(dsum + 2);
    GP_LAUNCH(pr_init_kernel, nblocks, 256, 0, stream, csr->deg, n, xa, inv_out);
    GP_LAUNCH(dangling_list_kernel, 1, 32, 0, stream, csr->deg, n, dang, n_dang);
    const double one_minus_alpha = 1 - alpha;
    bool converged = false;
    int it = 0;
    for (it = 1; it <= max_iter; ++it) {
        GP_LAUNCH(pr_contrib_kernel, nblocks + 1, 256, 0, stream, xa, inv_out, n, contrib, dang, n_dang, dsum);
        GP_LAUNCH(pr_pull_kernel, nblocks, 256, 0, stream, csr->rowptr_in, csr->col_in, contrib, xa, n, alpha,
                                                    one_minus_alpha, dsum, xb, (double *)b_partial.p);
        GP_LAUNCH(pr_err_kernel, 1, 256, 0, stream, (const double *)b_partial.p, nblocks, err);
        double h_err = 0.0;
        GP_CUDA_CHECK(cudaMemcpyAsync(&h_err, err, sizeof(double), cudaMemcpyDeviceToHost, stream));
        GP_CUDA_CHECK(cudaStreamSynchronize(stream));
        double *t = xa; xa = xb; xb = t;
        if (h_err < (double)n * tol) {
            converged = true;
            break;
        }
    }
    if (xa != d_x) GP_CUDA_CHECK(cudaMemcpyAsync(d_x, xa, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    GP_CUDA_CHECK(cudaStreamSynchronize(stream));
    int csr_err = 0;
    GP_CUDA_CHECK(cudaMemcpy(&csr_err, &csr->meta[GP_META_ERROR], sizeof(int), cudaMemcpyDeviceToHost));
    GP_REQUIRE(!(csr_err & GP_DEV_ERR_EDGE_RANGE), GP_ERR_INDEX_RANGE,
               "edge_index holds an entry outside [0, %lld)", n);
    if (iterations) *iterations = converged ? it : max_iter;
    GP_REQUIRE(converged, GP_ERR_NOT_CONVERGED, "pagerank: power iteration failed to converge in %d iterations",
               max_iter);
    return GP_OK;
}

// ------------------------------------------------------------------------------------------ eigenvector
// eigenvector_centrality (utils.py:44-48 -> nx.eigenvector_centrality_numpy): the eigenvector of A^T for
// its largest real eigenvalue (in-edge centrality), unit L2 norm, positive sum.  networkx hands the matrix to
// ARPACK; here: power iteration on (A^T + I) — the shift makes the Perron root strictly dominant also on
// periodic graphs — in float64, one thread per row with the in-neighbours added in ascending order, fixed-
// order reductions, until the L1 change of the normalised vector drops below N * tol.
namespace {

__global__ void ev_fill_kernel(double *x, long long n, double v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

__device__ __forceinline__ void block_partial(double v, double *partial)
{
    __shared__ double s_part[8];
    for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(FULL_MASK, v, m);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w];
        partial[blockIdx.x] = t;
    }
}

// y = (A^T + I) x; partial[b] = this block's share of |y|^2
__global__ void __launch_bounds__(256)
ev_pull_kernel(const int *__restrict__ rp_in, const int *__restrict__ col_in, const double *__restrict__ x, long long n,
               double *__restrict__ y, double *__restrict__ partial)
{
    double ss = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const int s = rp_in[v], e = rp_in[v + 1];
        double acc = x[v];
        int j = s;
        for (; j + 4 <= e; j += 4) {
            const double c0 = x[col_in[j]], c1 = x[col_in[j + 1]], c2 = x[col_in[j + 2]], c3 = x[col_in[j + 3]];
            acc = __dadd_rn(acc, c0);
            acc = __dadd_rn(acc, c1);
            acc = __dadd_rn(acc, c2);
            acc = __dadd_rn(acc, c3);
        }
        for (; j < e; ++j) acc = __dadd_rn(acc, x[col_in[j]]);
        y[v] = acc;
        ss += acc * acc;
    }
    block_partial(ss, partial);
}

// x_new = y / |y|; partial[b] = this block's share of |x_new - x|_1.  `ss` holds |y|^2.
__global__ void __launch_bounds__(256)
ev_scale_kernel(const double *__restrict__ y, const double *__restrict__ ss, double *__restrict__ x, long long n,
                double *__restrict__ partial)
{
    const double norm = sqrt(*ss);
    double err = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const double xv = __ddiv_rn(y[v], norm);
        err += fabs(xv - x[v]);
        x[v] = xv;
    }
    block_partial(err, partial);
}

}  // namespace

extern "C" int gp_eigenvector(const gp_csr_t *csr_, double tol, int32_t max_iter, double *d_x, int32_t *iterations,
                              gp_stream_t stream_)
{
    gp_csr *csr = const_cast<gp_csr *>(csr_);
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(csr != nullptr && d_x != nullptr, GP_ERR_INVALID, "gp_eigenvector: NULL argument");
    GP_REQUIRE(csr->built, GP_ERR_INVALID, "gp_eigenvector: the CSR has not been built");
    GP_REQUIRE(tol >= 0 && max_iter >= 1, GP_ERR_INVALID, "gp_eigenvector: bad tolerance / iteration limit");
    if (iterations) *iterations = 0;
    const long long n = csr->num_nodes;
    if (n == 0) return GP_OK;
    GP_TRY(gp_csr_ensure_in(csr, stream));
    const int nblocks = blocks_for(n);
    DevBuf b_y, b_partial, b_small;
    GP_CUDA_CHECK(cudaMalloc(&b_y.p, sizeof(double) * n));
    GP_CUDA_CHECK(cudaMalloc(&b_partial.p, sizeof(double) * nblocks));
    GP_CUDA_CHECK(cudaMalloc(&b_small.p, 64));
    double *y = (double *)b_y.p, *partial = (double *)b_partial.p, *scalar = (double *)b_small.p;
    GP_LAUNCH(ev_fill_kernel, nblocks, 256, 0, stream, d_x, n, 1.0 / sqrt((double)n));
    bool converged = false;
    int it = 0;
    const int check_every = 8;  // one host round trip per 8 iterations
    while (it < max_iter && !converged) {
        double h_err = 0.0;
        for (int q = 0; q < check_every && it < max_iter; ++q, ++it) {
            GP_LAUNCH(ev_pull_kernel, nblocks, 256, 0, stream, csr->rowptr_in, csr->col_in, d_x, n, y, partial);
            GP_LAUNCH(pr_err_kernel, 1, 256, 0, stream, (const double *)partial, nblocks, scalar);
            GP_LAUNCH(ev_scale_kernel, nblocks, 256, 0, stream, y, scalar, d_x, n, partial);
            GP_LAUNCH(pr_err_kernel, 1, 256, 0, stream, (const double *)partial, nblocks, scalar + 1);
        }
        GP_CUDA_CHECK(cudaMemcpyAsync(&h_err, scalar + 1, sizeof(double), cudaMemcpyDeviceToHost, stream));
        GP_CUDA_CHECK(cudaStreamSynchronize(stream));
        converged = h_err <= (double)n * tol;
    }
    int csr_err = 0;
    GP_CUDA_CHECK(cudaMemcpy(&csr_err, &csr->meta[GP_META_ERROR], sizeof(int), cudaMemcpyDeviceToHost));
    GP_REQUIRE(!(csr_err & GP_DEV_ERR_EDGE_RANGE), GP_ERR_INDEX_RANGE,
               "edge_index holds an entry outside [0, %lld)", n);
    if (iterations) *iterations = it;
    GP_REQUIRE(converged, GP_ERR_NOT_CONVERGED,
               "eigenvector centrality: power iteration did not reach tol %.3g in %d iterations", tol, max_iter);
    return GP_OK;
}

static int topk_common(const int *d_score_i32, const double *d_score_f64, int64_t n, int64_t k, int64_t *d_out,
                       cudaStream_t stream)
{
    GP_REQUIRE(n >= 0 && k >= 0, GP_ERR_INVALID, "gp_topk_stable: negative size");
    if (n == 0) return GP_OK;
    GP_REQUIRE(d_out != nullptr, GP_ERR_INVALID, "gp_topk_stable: NULL argument");
    // list[-k:] : k == 0 keeps everything (reference quirk), k > n keeps everything
    const int64_t take = (k == 0 || k > n) ? n : k;
    const bool f64 = d_score_f64 != nullptr;
    GpSortWorkspace ws;
    int rc = gp_sort_workspace_create(&ws, n, f64);
    DevBuf b_keys, b_vals;
    if (rc == GP_OK && cudaMalloc(&b_keys.p, sizeof(u64) * n) != cudaSuccess) rc = GP_ERR_OOM;
    if (rc == GP_OK && f64 && cudaMalloc(&b_vals.p, sizeof(u32) * n) != cudaSuccess) rc = GP_ERR_OOM;
    if (rc != GP_OK) {
        gp_sort_workspace_free(&ws);
        if (rc == GP_ERR_OOM) gp_set_error("gp_topk_stable: out of device memory");
        return rc;
    }
    u64 *keys = (u64 *)b_keys.p, *skeys = nullptr;
    u32 *vals = (u32 *)b_vals.p, *svals = nullptr;
    if (f64) {
        GP_LAUNCH(topk_keys_f64_kernel, blocks_for(n), 256, 0, stream, d_score_f64, n, keys, vals);
        rc = gp_radix_sort(&ws, keys, vals, nullptr, n, 0, 64, stream, &skeys, &svals);
    } else {
        GP_LAUNCH(topk_keys_i32_kernel, blocks_for(n), 256, 0, stream, d_score_i32, n, keys);
        rc = gp_radix_sort(&ws, keys, nullptr, nullptr, n, 32, 64, stream, &skeys, nullptr);
    }
    if (rc == GP_OK) {
        GP_LAUNCH(topk_emit_kernel, blocks_for(take), 256, 0, stream, skeys, f64 ? svals : nullptr, n, take, (long long *)d_out);
        if (cudaGetLastError() != cudaSuccess) rc = GP_ERR_CUDA;
    }
    cudaStreamSynchronize(stream);  // temporaries are freed below
    gp_sort_workspace_free(&ws);
    return rc;
}

extern "C" int gp_topk_stable_i32(const int32_t *d_score, int64_t num_nodes, int64_t k, int64_t *d_out,
                                  gp_stream_t stream)
{
    GP_REQUIRE(d_score != nullptr || num_nodes == 0, GP_ERR_INVALID, "gp_topk_stable_i32: NULL score");
    return topk_common(d_score, nullptr, num_nodes, k, d_out, (cudaStream_t)stream);
}

extern "C" int gp_topk_stable_f64(const double *d_score, int64_t num_nodes, int64_t k, int64_t *d_out,
                                  gp_stream_t stream)
{
    GP_REQUIRE(d_score != nullptr || num_nodes == 0, GP_ERR_INVALID, "gp_topk_stable_f64: NULL score");
    return topk_common(nullptr, d_score, num_nodes, k, d_out, (cudaStream_t)stream);
}
