// gp_betweenness.cu — betweenness-centrality scores for the 'betweenness_centrality' anchor sampler
// (reference utils.py:32-36: nx.betweenness_centrality(G), then the stable top-k of utils.py:35-36).
//
// networkx runs Brandes' algorithm once per source s (betweenness.py: _single_source_shortest_path_basic,
// _accumulate_basic, _rescale with normalized=True, endpoints=False on the DiGraph):
//   forward   sigma[s] = 1; BFS over out-edges; sigma[w] = sum of sigma[v] over predecessors v one hop closer
//   backward  in reverse BFS order: coeff = (1 + delta[w]) / sigma[w]; delta[v] += sigma[v] * coeff for every
//             predecessor v of w;  betweenness[w] += delta[w]  (w != s)
//   rescale   betweenness[v] *= 1 / ((N-1)(N-2))   (N > 2)
// = N BFS runs; the reference does them in Python (hours at Flickr size).  Here: 32*S sources per batch, S = 1, 2
// or 4 sources PER LANE.  Per-batch state is node-major — dist[N][32*S] (hop count per source, u8 or u16),
// sigma[N][32*S] and coeff[N][32*S] float64 — so a warp reads the hop counts of one neighbour for all sources of
// the batch with one coalesced request (one word per lane) and its path counts as one contiguous row.  Both sweeps are level-synchronous and *pull* (owner
// computes): forward, an unsettled row sums sigma over its in-neighbours settled one level earlier; backward,
// a row at level l sums sigma[v] * coeff[w] over its out-neighbours w at level l+1, where
// coeff[w] = (1 + delta[w]) / sigma[w] was stored when w was finalised.  No atomics touch floating-point
// data, every sum has a fixed order (neighbours ascending; hub rows: 64-edge chunks whose partial sums are
// added in chunk order by the last-arriving warp), so the scores are run-to-run deterministic.  They are NOT
// bit-equal to networkx: its delta sums run in reverse BFS-queue order, which depends on the column order of
// edge_index; sigma is exact (integer-valued), delta and the scores agree to a few ulp (tests: rtol 1e-10).
//
// One persistent cooperative kernel walks all levels of a group of batches with a grid barrier per level;
// work is dealt in tiles of items (row or 64-edge chunk of a row) from rotating ticket counters.
#include <chrono>
#include <vector>

#include "gp_internal.h"

namespace {

constexpr int BC_CHUNK = 64;    // edges per work item
constexpr int BC_TILE = 4;      // items per ticket (lane k of the warp holds the descriptor of item k)
constexpr int BC_THREADS = 256;
constexpr int BC_MINB = 3;
constexpr int BC_UNROLL = 8;    // neighbour hop-count loads in flight per lane
// Words of the sync block.  The barrier word is polled by every CTA while late warps still draw tickets, so the
// barrier, each ticket counter and each flag live in their own 128-byte line.
constexpr int BC_SYNC_BARRIER = 0;
constexpr int BC_SYNC_TICKET = 32;    // + 32 * (sweep % 3)
constexpr int BC_SYNC_FLAG = 128;     // + 32 * (sweep % 3)
constexpr int BC_SYNC_OVERFLOW = 224;
constexpr int BC_SYNC_SWEEPS = 225;
constexpr int BC_SYNC_WORDS = 256;

struct BcList {
    const int4 *items;       // {row, first edge, count | chunk << 8, hub slot or -1}
    int num_items;
    const int *col;          // column array the items index
    const int *hub_chunks;   // [hubs] chunks of each chunked row
    const int *hub_first;    // [hubs] first partial slot of each chunked row
};

struct BcParams {
    BcList fwd, bwd;         // in-edge items (forward sweep), out-edge items (backward sweep)
    long long n;
    void *dist;              // [n][32 * S] hop count from each source of the batch (all ones = not reached)
    double *sigma;           // [n][32 * S] shortest-path counts
    double *coeff;           // [n][32 * S] (1 + delta) / sigma of finalised rows
    double *partial;         // [chunks of chunked rows][32 * S]
    u32 *arrive;             // [hubs] arrival counters (left at zero by the finaliser)
    double *bc;              // [n] running sum of delta over the sources done so far
    u32 *sync;               // BC_SYNC_* words: barrier, 3 ticket counters, 3 "something settled" flags, overflow
    long long batch0, batch1;
};

__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = __dadd_rn(v, __shfl_xor_sync(FULL_MASK, v, m));
    return v;
}

// A lane's S hop counts of one row as one aligned load / store (L2-coherent load: other SMs write these words).
template <int BYTES>
__device__ __forceinline__ u64 ld_state(const void *p)
{
    if (BYTES == 1) return __ldcg(reinterpret_cast<const unsigned char *>(p));
    if (BYTES == 2) return __ldcg(reinterpret_cast<const unsigned short *>(p));
    if (BYTES == 4) return __ldcg(reinterpret_cast<const unsigned int *>(p));
    return __ldcg(reinterpret_cast<const unsigned long long *>(p));
}

template <int BYTES>
__device__ __forceinline__ void st_state(void *p, u64 v)
{
    if (BYTES == 1) *reinterpret_cast<unsigned char *>(p) = (unsigned char)v;
    else if (BYTES == 2) *reinterpret_cast<unsigned short *>(p) = (unsigned short)v;
    else if (BYTES == 4) *reinterpret_cast<unsigned int *>(p) = (unsigned int)v;
    else *reinterpret_cast<unsigned long long *>(p) = v;
}

// One sweep over a work list.  fwd: settle level `lvl + 1` from level `lvl`.  !fwd: finalise level `lvl`
// from level `lvl + 1` (`maxl` = deepest level of the batch, whose rows have delta = 0).
// A lane owns S sources: its S hop counts of a row are one word, its S sigma / coeff values one 8*S-byte run.
// ONE copy of this code serves both directions and every item of a tile (the direction is a uniform run-time
// flag, the item loop is rolled with the next item's row state and columns prefetched): unrolled four items
// deep and instantiated per direction it was 105 KB of SASS and the kernel stalled on instruction fetch
// (ncu: stalled_no_instruction 8.7 per issue, the same as the first cdist kernel).
template <typename DT, int S>
__device__ __forceinline__ void bc_sweep(const BcParams &p, const BcList &L, bool fwd, int lvl, int maxl, u32 *ticket,
                                         u32 *flag)
{
    constexpr DT INF = (DT)~(DT)0;
    constexpr int BYTES = S * (int)sizeof(DT);
    constexpr int BITS = 8 * (int)sizeof(DT);
    constexpr int W = 32 * S;  // sources per batch
    const int lane = threadIdx.x & 31;
    char *dist = reinterpret_cast<char *>(p.dist);
    const DT want = (DT)(fwd ? lvl : lvl + 1);    // hop count a contributing neighbour must have
    const DT own = fwd ? INF : (DT)lvl;           // hop count of the (row, source) pairs this sweep finalises
    const bool leaf = !fwd && (lvl + 1 == maxl);  // neighbours at the deepest level: coeff = 1 / sigma
    const double *nbr_val = (fwd || leaf) ? p.sigma : p.coeff;  // what is read per contributing neighbour
    bool found_any = false;
    const int tiles = (L.num_items + BC_TILE - 1) / BC_TILE;
    // first tile of a warp: its own index (no atomic); further tiles are drawn from the ticket counter, each
    // fetched while the previous tile is processed
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    u32 next = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (;;) {
        const u32 t = __shfl_sync(FULL_MASK, next, 0);
        if (t >= (u32)tiles) break;
        if (lane == 0) next = nwarps + atomicAdd(ticket, 1u);
        const int i0 = (int)t * BC_TILE;
        const int nit = min(BC_TILE, L.num_items - i0);
        // round trip 1: the tile's descriptors (lane k holds item k); round trip 2: every item's row state and
        // first 32 columns, issued together, so that items with nothing to do this sweep cost no round trip each
        int4 mine_it = make_int4(0, 0, 0, -1);
        if (lane < nit) mine_it = __ldg(L.items + i0 + lane);
        u64 raw_t[BC_TILE];
        int col_t[BC_TILE];
#pragma unroll
        for (int k = 0; k < BC_TILE; ++k) {
            const int r = __shfl_sync(FULL_MASK, mine_it.x, k), b = __shfl_sync(FULL_MASK, mine_it.y, k);
            const int c = __shfl_sync(FULL_MASK, mine_it.z, k) & 0xFF;
            raw_t[k] = ld_state<BYTES>(dist + ((size_t)r * 32 + lane) * BYTES);
            col_t[k] = lane < c ? __ldg(L.col + b + lane) : r;
        }
#pragma unroll 1
        for (int k = 0; k < nit; ++k) {
            const int row = __shfl_sync(FULL_MASK, mine_it.x, k), beg = __shfl_sync(FULL_MASK, mine_it.y, k);
            const int z = __shfl_sync(FULL_MASK, mine_it.z, k), hub = __shfl_sync(FULL_MASK, mine_it.w, k);
            u64 rowraw = raw_t[0];
            int col0 = col_t[0];
#pragma unroll
            for (int q = 1; q < BC_TILE; ++q) {  // k is warp-uniform: a select chain, not a local-memory array
                if (k == q) {
                    rowraw = raw_t[q];
                    col0 = col_t[q];
                }
            }
            do {
                const int cnt = z & 0xFF, chunk = z >> 8;
                u32 mine = 0;  // bit s: source s of this lane is finalised by this sweep
#pragma unroll
                for (int s = 0; s < S; ++s) mine |= ((DT)(rowraw >> (s * BITS)) == own ? 1u : 0u) << s;
                if (!__any_sync(FULL_MASK, mine != 0)) break;
                const size_t rbase = ((size_t)row * 32 + lane) * S;
                double sv[S], acc[S];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    acc[s] = 0.0;
                    sv[s] = (!fwd && ((mine >> s) & 1u)) ? __ldcg(p.sigma + rbase + s) : 1.0;
                }
                for (int base = 0; base < cnt; base += 32) {
                    int c = col0;
                    if (base > 0) c = (base + lane < cnt) ? __ldg(L.col + beg + base + lane) : row;
                    const int m = min(32, cnt - base);
                    for (int j0 = 0; j0 < m; j0 += BC_UNROLL) {
                        int v[BC_UNROLL];
                        u64 dv[BC_UNROLL];
#pragma unroll
                        for (int q = 0; q < BC_UNROLL; ++q) {
                            // padding slots re-read the row itself, whose own hop counts never equal `want`
                            v[q] = __shfl_sync(FULL_MASK, c, (j0 + q) & 31);
                            if (j0 + q >= m) v[q] = row;
                        }
#pragma unroll
                        for (int q = 0; q < BC_UNROLL; ++q)
                            dv[q] = ld_state<BYTES>(dist + ((size_t)v[q] * 32 + lane) * BYTES);
#pragma unroll
                        for (int q = 0; q < BC_UNROLL; ++q) {
                            const size_t vb = ((size_t)v[q] * 32 + lane) * S;
#pragma unroll
                            for (int s = 0; s < S; ++s) {
                                if (((mine >> s) & 1u) && (DT)(dv[q] >> (s * BITS)) == want) {
                                    // forward: + sigma[v]; backward: + sigma[row] * coeff[v] (a leaf's coeff is
                                    // 1 / sigma[v]); sv is 1.0 forward and x * 1.0 is exact
                                    double cw = __ldcg(nbr_val + vb + s);
                                    if (leaf) cw = __ddiv_rn(1.0, cw);
                                    acc[s] = __dadd_rn(acc[s], fwd ? cw : __dmul_rn(sv[s], cw));
                                }
                            }
                        }
                    }
                }
                if (hub >= 0) {
                    // chunk of a long row: park the partial sums; the warp that arrives last adds them in chunk order
                    const int chunks = __ldg(L.hub_chunks + hub), first = __ldg(L.hub_first + hub);
                    double *slot = p.partial + ((size_t)first + chunk) * W + (size_t)lane * S;
#pragma unroll
                    for (int s = 0; s < S; ++s) slot[s] = acc[s];
                    __syncwarp();
                    u32 old = 0;
                    if (lane == 0) {
                        __threadfence();  // the warp's partial sums before the arrival
                        old = atomicAdd(p.arrive + hub, 1u);
                    }
                    old = __shfl_sync(FULL_MASK, old, 0);
                    if (old != (u32)(chunks - 1)) break;
                    __threadfence();
                    if (lane == 0) p.arrive[hub] = 0;
#pragma unroll
                    for (int s = 0; s < S; ++s) acc[s] = 0.0;
                    const double *part = p.partial + (size_t)first * W + (size_t)lane * S;
                    constexpr int PB = 16 / S;  // 16 loads in flight, added in chunk order
                    for (int ch0 = 0; ch0 < chunks; ch0 += PB) {
                        double pv[PB][S];
#pragma unroll
                        for (int q = 0; q < PB; ++q)
#pragma unroll
                            for (int s = 0; s < S; ++s)
                                pv[q][s] = ch0 + q < chunks ? __ldcg(part + (size_t)(ch0 + q) * W + s) : 0.0;
#pragma unroll
                        for (int q = 0; q < PB; ++q)
                            if (ch0 + q < chunks) {
#pragma unroll
                                for (int s = 0; s < S; ++s) acc[s] = __dadd_rn(acc[s], pv[q][s]);
                            }
                    }
                }
                if (fwd) {
                    u64 raw = rowraw;
                    bool changed = false;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        // sigma >= 1 for every settled node, so a non-zero sum means "has a parent at level lvl"
                        if (((mine >> s) & 1u) && acc[s] != 0.0) {
                            p.sigma[rbase + s] = acc[s];
                            raw = (raw & ~((u64)INF << (s * BITS))) | ((u64)(DT)(lvl + 1) << (s * BITS));
                            changed = true;
                        }
                    }
                    if (changed) {  // the lane owns the whole word: readers see the old or the new hop counts
                        st_state<BYTES>(dist + ((size_t)row * 32 + lane) * BYTES, raw);
                        found_any = true;
                    }
                } else {
                    double mysum = 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if ((mine >> s) & 1u) {
                            p.coeff[rbase + s] = __ddiv_rn(__dadd_rn(1.0, acc[s]), sv[s]);
                            mysum = __dadd_rn(mysum, acc[s]);
                        }
                    }
                    const double tot = warp_sum_f64(mysum);
                    if (lane == 0) p.bc[row] = __dadd_rn(__ldcg(p.bc + row), tot);  // the row has one owner per sweep
                }
            } while (0);
        }
    }
    if (fwd && __any_sync(FULL_MASK, found_any) && lane == 0) st_relaxed_u32(flag, 1u);
}

template <typename DT, int S>
__global__ void __launch_bounds__(BC_THREADS, (S == 4 ? 2 : BC_MINB)) bc_kernel(BcParams p)
{
    constexpr DT INF = (DT)~(DT)0;
    constexpr int BYTES = S * (int)sizeof(DT);
    constexpr int BITS = 8 * (int)sizeof(DT);
    constexpr int W = 32 * S;
    const int lane = threadIdx.x & 31;
    const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    char *dist = reinterpret_cast<char *>(p.dist);
    u32 target = 0;  // barrier arrivals so far (the counter is zeroed before every launch)
    u32 pass = 0;    // sweeps so far: ticket counter / flag `pass % 3` is live, `(pass + 1) % 3` is being cleared
    bool overflow = false;
    for (long long batch = p.batch0; batch < p.batch1 && !overflow; ++batch) {
        const long long src0 = batch * W + (long long)lane * S;  // this lane's first source (>= n: idle)
        for (long long w = gwarp; w < p.n; w += nwarps) {
            u64 raw = ~0ull;
            const long long rel = w - src0;
            if (rel >= 0 && rel < S) {
                raw &= ~((u64)INF << (rel * BITS));
                p.sigma[((size_t)w * 32 + lane) * S + rel] = 1.0;
            }
            st_state<BYTES>(dist + ((size_t)w * 32 + lane) * BYTES, raw);
        }
        grid_barrier(p.sync + BC_SYNC_BARRIER, target, gridDim.x);
        // forward sweeps lvl = 0, 1, ... until one settles nothing (maxl = deepest level), then backward sweeps
        // lvl = maxl - 1 ... 1 (rows at maxl have delta = 0, the sources at level 0 are not scored)
        bool fwd = true;
        int lvl = 0, maxl = 0;
        for (;;) {
            bc_sweep<DT, S>(p, fwd ? p.fwd : p.bwd, fwd, lvl, maxl, p.sync + BC_SYNC_TICKET + 32 * (pass % 3),
                            p.sync + BC_SYNC_FLAG + 32 * (pass % 3));
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                // counter / flag of the NEXT sweep: last touched two sweeps ago, i.e. before the previous barrier
                p.sync[BC_SYNC_TICKET + 32 * ((pass + 1) % 3)] = 0;
                p.sync[BC_SYNC_FLAG + 32 * ((pass + 1) % 3)] = 0;
            }
            grid_barrier(p.sync + BC_SYNC_BARRIER, target, gridDim.x);
            const u32 found = ld_relaxed_u32(p.sync + BC_SYNC_FLAG + 32 * (pass % 3));
            ++pass;
            if (fwd) {
                if (found) {
                    ++lvl;
                    if (lvl + 1 >= (int)INF) {  // the next level could not be told from "not reached"
                        overflow = true;
                        break;
                    }
                    continue;
                }
                maxl = lvl;  // levels 0..maxl exist
                fwd = false;
            }
            if (--lvl < 1) break;
        }
        if (overflow && blockIdx.x == 0 && threadIdx.x == 0) p.sync[BC_SYNC_OVERFLOW] = 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) p.sync[BC_SYNC_SWEEPS] = pass;  // sweeps of this launch (diagnostics)
}

__global__ void bc_scale_kernel(double *bc, long long n, double scale)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bc[i] = __dmul_rn(bc[i], scale);
}

// Work list of one direction, built on the host from the row extents (setup only: N integers each way).
struct HostList {
    std::vector<int4> items;
    std::vector<int> hub_chunks, hub_first;
    int partial_slots = 0;
};

// keep_empty: rows without neighbours still get an item (backward sweep: their coeff = 1 / sigma must be
// written for their predecessors to read; forward, a row without in-edges is never settled and is skipped).
void build_list(const std::vector<int> &first, const std::vector<int> &count, bool keep_empty, HostList &out)
{
    const size_t n = count.size();
    out.items.reserve(n + n / 8);
    // chunked (long) rows first: their chunks are spread over many warps and the row is finalised by whichever
    // arrives last, so started late they would be the tail every other CTA waits for at the level barrier
    for (int pass = 0; pass < 2; ++pass)
    for (size_t r = 0; r < n; ++r) {
        const int d = count[r];
        if (d == 0 && !keep_empty) continue;
        if ((d > BC_CHUNK) != (pass == 0)) continue;
        if (d <= BC_CHUNK) {
            out.items.push_back(make_int4((int)r, first[r], d, -1));
            continue;
        }
        const int chunks = (d + BC_CHUNK - 1) / BC_CHUNK;
        const int hub = (int)out.hub_chunks.size();
        out.hub_chunks.push_back(chunks);
        out.hub_first.push_back(out.partial_slots);
        out.partial_slots += chunks;
        for (int c = 0; c < chunks; ++c) {
            const int cnt = (c + 1 < chunks) ? BC_CHUNK : d - c * BC_CHUNK;
            out.items.push_back(make_int4((int)r, first[r] + c * BC_CHUNK, cnt | (c << 8), hub));
        }
    }
}

// Uploads one work list into the handle's scratch slots [slot0, slot0 + 3).
int upload_list(gp_csr *csr, int slot0, const HostList &h, BcList &out, const int *col, cudaStream_t stream)
{
    void *items = nullptr, *hub_chunks = nullptr, *hub_first = nullptr;
    GP_TRY(gp_csr_scratch(csr, slot0, h.items.size() * sizeof(int4), &items));
    GP_TRY(gp_csr_scratch(csr, slot0 + 1, h.hub_chunks.size() * sizeof(int), &hub_chunks));
    GP_TRY(gp_csr_scratch(csr, slot0 + 2, h.hub_first.size() * sizeof(int), &hub_first));
    if (!h.items.empty())
        GP_CUDA_CHECK(cudaMemcpyAsync(items, h.items.data(), h.items.size() * sizeof(int4), cudaMemcpyHostToDevice, stream));
    if (!h.hub_chunks.empty()) {
        GP_CUDA_CHECK(cudaMemcpyAsync(hub_chunks, h.hub_chunks.data(), h.hub_chunks.size() * sizeof(int),
                                      cudaMemcpyHostToDevice, stream));
        GP_CUDA_CHECK(cudaMemcpyAsync(hub_first, h.hub_first.data(), h.hub_first.size() * sizeof(int),
                                      cudaMemcpyHostToDevice, stream));
    }
    out.items = (const int4 *)items;
    out.num_items = (int)h.items.size();
    out.col = col;
    out.hub_chunks = (const int *)hub_chunks;
    out.hub_first = (const int *)hub_first;
    return GP_OK;
}

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename DT, int S>
int run_batches(const BcParams &base, long long n, int group, u32 *h_overflow, cudaStream_t stream)
{
    const long long batches = (n + 32 * S - 1) / (32 * S);
    int occ = 0;
    GP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bc_kernel<DT, S>, BC_THREADS, 0));
    GP_REQUIRE(occ >= 1, GP_ERR_CUDA, "betweenness kernel does not fit on an SM");
    if (occ > (S == 4 ? 2 : BC_MINB)) occ = (S == 4 ? 2 : BC_MINB);
    const int blocks = occ * gp_sm_count();
    *h_overflow = 0;
    const bool trace = getenv("GP_BC_TRACE") != nullptr;
    for (long long b0 = 0; b0 < batches && !*h_overflow; b0 += group) {
        const double t0 = trace ? now_ms() : 0.0;
        BcParams p = base;
        p.batch0 = b0;
        p.batch1 = b0 + group < batches ? b0 + group : batches;
        GP_CUDA_CHECK(cudaMemsetAsync(p.sync, 0, BC_SYNC_WORDS * sizeof(u32), stream));
        void *args[] = {&p};
        gp_count_launch();
        GP_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)bc_kernel<DT, S>, dim3(blocks), dim3(BC_THREADS), args, 0,
                                                  stream));
        GP_CUDA_CHECK(cudaMemcpyAsync(h_overflow, p.sync + BC_SYNC_OVERFLOW, sizeof(u32), cudaMemcpyDeviceToHost, stream));
        GP_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (trace) {
            u32 sweeps = 0;
            GP_CUDA_CHECK(cudaMemcpy(&sweeps, p.sync + BC_SYNC_SWEEPS, sizeof(u32), cudaMemcpyDeviceToHost));
            const double t1 = now_ms();
            fprintf(stderr, "[gp_betweenness] S=%d bytes=%d blocks=%d batches %lld..%lld: %u sweeps, %.3f ms (%.1f us per sweep)\n",
                    S, (int)sizeof(DT), blocks, b0, p.batch1, sweeps, t1 - t0, sweeps ? (t1 - t0) * 1e3 / sweeps : 0.0);
        }
    }
    return GP_OK;
}

}  // namespace

extern "C" int gp_betweenness(const gp_csr_t *csr_, double *d_score, gp_stream_t stream_)
{
    gp_csr *csr = const_cast<gp_csr *>(csr_);
    cudaStream_t stream = (cudaStream_t)stream_;
    GP_REQUIRE(csr != nullptr && d_score != nullptr, GP_ERR_INVALID, "gp_betweenness: NULL argument");
    GP_REQUIRE(csr->built, GP_ERR_INVALID, "gp_betweenness: the CSR has not been built");
    const long long n = csr->num_nodes;
    if (n == 0) return GP_OK;
    GP_CUDA_CHECK(cudaMemsetAsync(d_score, 0, (size_t)n * sizeof(double), stream));
    if (n <= 2) {  // no node can lie strictly inside a path (and networkx skips the rescale)
        GP_CUDA_CHECK(cudaStreamSynchronize(stream));
        return GP_OK;
    }
    const bool trace = getenv("GP_BC_TRACE") != nullptr;
    const double t_begin = now_ms();
    GP_TRY(gp_csr_ensure_in(csr, stream));

    // ---- work lists (host, from the row extents)
    std::vector<int> first_in((size_t)n + 1), cnt_in((size_t)n), first_out((size_t)n), cnt_out((size_t)n);
    GP_CUDA_CHECK(cudaMemcpyAsync(first_in.data(), csr->rowptr_in, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, stream));
    GP_CUDA_CHECK(cudaMemcpyAsync(first_out.data(), csr->row_start, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    GP_CUDA_CHECK(cudaMemcpyAsync(cnt_out.data(), csr->deg, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    GP_CUDA_CHECK(cudaStreamSynchronize(stream));
    for (long long r = 0; r < n; ++r) cnt_in[(size_t)r] = first_in[(size_t)r + 1] - first_in[(size_t)r];
    first_in.resize((size_t)n);
    HostList hf, hb;
    build_list(first_in, cnt_in, false, hf);
    build_list(first_out, cnt_out, true, hb);

    // the pageable host vectors above must outlive the asynchronous uploads: the stream is synchronised by the
    // first launch group below before they go out of scope
    BcParams p;
    memset(&p, 0, sizeof(p));
    GP_TRY(upload_list(csr, 0, hf, p.fwd, csr->col_in, stream));
    GP_TRY(upload_list(csr, 3, hb, p.bwd, csr->col, stream));
    p.n = n;
    p.bc = d_score;

    // sources per lane (1, 2 or 4 -> 32, 64 or 128 sources per batch): more sources per batch amortise the
    // per-level barrier and list scan; the work per source stays the same (the sweeps are bound by instruction
    // issue), so the gain flattens: Flickr-shape 2.8 s for 1, 2 and 4.  GP_BC_SOURCES overrides.
    int spl = n >= 512 ? 2 : 1;
    if (const char *e = getenv("GP_BC_SOURCES")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4) spl = v;
    }
    // workspace in the handle's scratch slots: allocating and freeing ~100 MB per call made repeated calls in one
    // process up to 2.4x slower (the time went into cudaFree / cudaMalloc, not into the sweeps)
    const size_t cells = (size_t)n * 32 * spl;
    const size_t slots = (size_t)(hf.partial_slots > hb.partial_slots ? hf.partial_slots : hb.partial_slots);
    const size_t hubs = hf.hub_chunks.size() > hb.hub_chunks.size() ? hf.hub_chunks.size() : hb.hub_chunks.size();
    void *arrive = nullptr;
    GP_TRY(gp_csr_scratch(csr, 6, cells * sizeof(uint16_t), &p.dist));
    GP_TRY(gp_csr_scratch(csr, 7, cells * sizeof(double), (void **)&p.sigma));
    GP_TRY(gp_csr_scratch(csr, 8, cells * sizeof(double), (void **)&p.coeff));
    GP_TRY(gp_csr_scratch(csr, 9, (slots ? slots : 1) * 32 * spl * sizeof(double), (void **)&p.partial));
    GP_TRY(gp_csr_scratch(csr, 10, (hubs ? hubs : 1) * sizeof(u32), &arrive));
    GP_TRY(gp_csr_scratch(csr, 11, BC_SYNC_WORDS * sizeof(u32), (void **)&p.sync));
    GP_CUDA_CHECK(cudaMemsetAsync(arrive, 0, (hubs ? hubs : 1) * sizeof(u32), stream));
    p.arrive = (u32 *)arrive;

    const double t_setup = now_ms();
    int group = 64;  // batches per cooperative launch: bounds the run time of one launch
    if (const char *e = getenv("GP_BC_GROUP")) group = atoi(e) > 0 ? atoi(e) : group;
    u32 overflow = 0;
    // 8-bit hop counts first (one sector per neighbour and 32 sources); a graph deeper than 253 hops restarts
    // with 16 bits
    if (spl == 4) GP_TRY((run_batches<uint8_t, 4>(p, n, group, &overflow, stream)));
    else if (spl == 2) GP_TRY((run_batches<uint8_t, 2>(p, n, group, &overflow, stream)));
    else GP_TRY((run_batches<uint8_t, 1>(p, n, group, &overflow, stream)));
    if (overflow) {
        GP_CUDA_CHECK(cudaMemsetAsync(d_score, 0, (size_t)n * sizeof(double), stream));
        GP_CUDA_CHECK(cudaMemsetAsync(arrive, 0, (hubs ? hubs : 1) * sizeof(u32), stream));
        if (spl == 4) GP_TRY((run_batches<uint16_t, 4>(p, n, group, &overflow, stream)));
        else if (spl == 2) GP_TRY((run_batches<uint16_t, 2>(p, n, group, &overflow, stream)));
        else GP_TRY((run_batches<uint16_t, 1>(p, n, group, &overflow, stream)));
        GP_REQUIRE(!overflow, GP_ERR_LEVEL_OVERFLOW, "gp_betweenness: a hop distance >= 65534");
    }
    const double t_sweeps = now_ms();
    if (trace)
        fprintf(stderr, "[gp_betweenness] setup (in-edge csr, work lists, allocations) %.3f ms, sweeps %.3f ms\n",
                t_setup - t_begin, t_sweeps - t_setup);
    // _rescale(normalized=True, directed=True, endpoints=False): scale = 1 / ((N-1) * (N-2)), Python float ops
    const double scale = 1.0 / ((double)(n - 1) * (double)(n - 2));
    if (scale != 1.0) {
        GP_LAUNCH(bc_scale_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, d_score, n, scale);
        GP_CUDA_CHECK(cudaGetLastError());
    }
    GP_CUDA_CHECK(cudaStreamSynchronize(stream));
    return GP_OK;
}
