// gp_msbfs.cu — multi-source BFS over bit lanes, one persistent cooperative kernel (sm_100a).
//
// Replaces the N*K nx.shortest_path calls of the reference hot loop (utils.py:69-77, fanned
// out by utils.py:92-114): hops(node -> anchor_j) for all nodes and anchors is computed as a
// level-synchronous BFS from all anchors at once, 64 anchors per uint64 lane word.
//
// Formulation.  Paths run node -> anchor (row = source, utils.py:73), so the frontier moves
// against the edge direction.  In PULL form a row u ORs the frontier words of its
// OUT-neighbours: new[u] = (OR_{u->v} frontier[v]) & ~seen[u].  Each row is finalised by exactly
// one thread group, so there are no atomics on the lane state and a level needs one grid barrier.
//
// Work decomposition.  On the named graphs the lane state (a few MB) lives in L2 and a level is
// LATENCY bound (an SM has ~6K neighbour gathers per level, i.e. ~12 per thread pair), so the
// kernel is organised to keep dependent-load chains short and every gather of a chain in flight
// at once: the CSR builder (gp_csr.cu) emits a degree-ordered list of 16-byte row descriptors;
// a row of degree d is served by G = 1,2,4,8,16 "pair slots" of <= 8 edges each (G*8 >= d), hub
// rows are cut into 128-edge chunks (G = 16) whose partial ORs meet in a small accumulator that
// the last-arriving chunk finalises.  One slot = descriptor -> <=8 column indices -> <=8 frontier
// rows (all issued back to back) -> shuffle-OR over the G slots of the row -> finalise.
//
// Data layout (all uint64, node-major so one neighbour gather is one 32-byte sector at wb == 4):
//   result block R, 32 arrays of [batches][N][wb]:
//     R[0]        reached mask ("seen")
//     R[l]        l = 1..15: lanes FIRST reached at hop l.  These are simply the frontiers: level l
//                 writes its next-frontier into R[l] and nothing ever overwrites it, so recording
//                 the hop count costs no extra traffic at all on shallow graphs.
//     R[16 + q]   q = 0..15: bit q of the hop count of lanes first reached at hop >= 16 (deep
//                 graphs only; OR-ed in with fire-and-forget reductions at L2).
//   plus the seed frontier and two ping-pong frontiers used from hop 16 on.
// The uint16 / fp32 matrices are never scattered to: the epilogue (gp_epilogue.cu) decodes R row
// by row with fully coalesced stores.
#include "gp_msbfs.cuh"

#include <new>

namespace {

struct BfsParams {
    int n;
    int batches;
    int num_anchors;
    int hub_capacity;
    const int *__restrict__ col;
    const int4 *__restrict__ desc;
    const int *__restrict__ meta;   // csr meta words (class bases)
    const long long *__restrict__ anchors;
    u64 *result;                    // R[0..31], see the layout note above
    u64 *seeds;                     // level-0 frontier (the anchors)
    u64 *fr_a;                      // ping-pong frontiers for hops >= 16
    u64 *fr_b;
    long long plane_stride;         // words per array = batches * n * wb
    u64 *live;                      // [3][GP_BFS_MAX_LANE_WORDS]
    u64 *hub_acc;                   // [batches][hub_capacity][wb] partial ORs of hub rows
    u32 *hub_cnt;                   // [batches][hub_capacity] chunks arrived
    u64 *bar;                       // [2] grid barrier words (arrivals | any-count << 32), by level parity
    int *status;
    u64 *counters;
    u64 *trace;                     // optional per-warp phase clocks (diagnostics), else nullptr
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

__device__ __forceinline__ void red_or_u64(u64 *p, u64 v)
{
    asm volatile("red.relaxed.gpu.global.or.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u32 atom_add_acq_rel_u32(u32 *p, u32 v)
{
    u32 old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Grid barrier that also tells every CTA whether ANY CTA reached a new lane this level.
// Word layout: low 32 bits = arrivals (monotone), high 32 bits = CTAs that reported `any`.
// Two words alternate by level parity, so a CTA that is already one barrier ahead never
// pollutes the word slower CTAs are still polling.
__device__ __forceinline__ bool grid_barrier_any(u64 *bar_word, u32 &target, u32 &prev_any, u32 nblocks,
                                                 bool cta_any, u32 *s_bcast, int *s_any_flag)
{
    target += nblocks;
    __syncthreads();
    if (threadIdx.x == 0) {
        *s_any_flag = 0;  // every thread has read it (before the sync above); next writes come after the sync below
        const u64 inc = 1ull | ((u64)(cta_any ? 1u : 0u) << 32);
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar_word), "l"(inc) : "memory");
        u64 v;
        do {
            v = ld_relaxed_u64(bar_word);  // relaxed polling: an acquire load would flush this SM's L1 each time
        } while ((u32)v < target);
        u32 dummy;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(dummy) : "l"(bar_word) : "memory");
        *s_bcast = (u32)(v >> 32);
    }
    __syncthreads();
    const u32 total_any = *s_bcast;
    const bool any = total_any != prev_any;
    prev_any = total_any;
    return any;
}

#define GP_TRACE(slot)                                                                             \
    do {                                                                                           \
        if (p.trace != nullptr && lane == 0 && level <= 32)                                        \
            p.trace[(((size_t)(level - 1) * total_warps + gwarp) << 2) + (slot)] = (u64)clock64(); \
    } while (0)

template <int VW>
struct LevelCtx {
    const u64 *cur;
    u64 *nxt;
    u64 *seen;
    u64 *planes;   // R[16]: deep-hop bit planes
    long long plane_stride;
    int level;
    int zero_first, zero_count;  // bit planes [zero_first, zero_first + zero_count) are cleared during this sweep
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
    vstore<VW>(c.nxt + off, nw);  // for level <= 15 this IS the record "first reached at hop level"
    if (c.zero_count > 0) {
        u64 z[VW];
#pragma unroll
        for (int i = 0; i < VW; ++i) z[i] = 0;
        for (int q = c.zero_first; q < c.zero_first + c.zero_count; ++q)
            vstore<VW>(c.planes + (size_t)q * c.plane_stride + off, z);
    }
    if (any) {
#pragma unroll
        for (int i = 0; i < VW; ++i) {
            seenv[i] |= nw[i];
            live_acc[i] |= nw[i];
        }
        vstore<VW>(c.seen + off, seenv);
        if (c.level > GP_BFS_LEVEL_ARRAYS) {
            // deep graphs: OR the new lanes into every bit plane set in `level` (fire-and-forget at L2)
            for (int lb = c.level; lb; lb &= lb - 1) {
                u64 *pp = c.planes + (size_t)(__ffs(lb) - 1) * c.plane_stride + off;
#pragma unroll
                for (int i = 0; i < VW; ++i)
                    if (nw[i]) red_or_u64(pp + i, nw[i]);
            }
        }
    }
}

// WB lane words per node row; TPE threads share one edge (each loads VW = WB/TPE words).
template <int WB, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) msbfs_kernel(BfsParams p)
{
    constexpr int TPE = (WB == 4) ? 2 : 1;
    constexpr int VW = WB / TPE;
    constexpr int PPW = 32 / TPE;  // pair slots per warp iteration
    constexpr int WARPS = NT / 32;
    constexpr u32 LEADER_MASK = (TPE == 2) ? 0x3u : 0x1u;

    constexpr int PAIRS_CTA = NT / TPE;
    // Work cache: a warp visits the same pair slots every level, so the descriptors and column
    // indices of its first GP_BFS_CACHE_ITERS warp-iterations are loaded ONCE into shared memory
    // (48 B per slot); a level then costs a single dependent round trip (gathers + seen together).
    extern __shared__ __align__(16) unsigned char s_dyn[];
    int4 *s_desc = reinterpret_cast<int4 *>(s_dyn);                                        // [ITERS][PAIRS_CTA]
    int *s_col = reinterpret_cast<int *>(s_desc + GP_BFS_CACHE_ITERS * PAIRS_CTA);        // [ITERS][8][PAIRS_CTA]
    unsigned char *s_done = reinterpret_cast<unsigned char *>(s_col + GP_BFS_CACHE_ITERS * GP_SLOT_EDGES * PAIRS_CTA);
                                                                                           // [DONE_B][ITERS][PAIRS_CTA]
    __shared__ u32 s_live32[GP_BFS_MAX_LANE_WORDS * 2];
    __shared__ int s_ent_base[GP_NUM_CLASSES + 1];
    __shared__ int s_slot_base[GP_NUM_CLASSES + 1];
    __shared__ u32 s_bcast;
    __shared__ int s_any;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = (TPE == 2) ? (lane & 1) : 0;
    const int woff = half * VW;
    const int pairlane = lane / TPE;
    const long long gthreads = (long long)gridDim.x * NT;
    const long long gtid = (long long)blockIdx.x * NT + tid;
    const int gwarp = blockIdx.x * WARPS + warp, total_warps = gridDim.x * WARPS;
    const int n = p.n, lw = p.batches * WB;
    u32 bar_target[2] = {0, 0}, bar_prev_any[2] = {0, 0};
    u64 gathers = 0;

    for (int i = tid; i < GP_BFS_MAX_LANE_WORDS * 2; i += NT) s_live32[i] = 0;
    if (tid <= GP_NUM_CLASSES) {
        s_ent_base[tid] = p.meta[GP_META_ENT_BASE + tid];
        s_slot_base[tid] = p.meta[GP_META_SLOT_BASE + tid];
    }
    if (tid == 0) s_any = 0;

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
        atomicOr(p.result + off, bit);
        atomicOr(p.seeds + off, bit);
        atomicOr(p.live + 1 * GP_BFS_MAX_LANE_WORDS + b * WB + w, bit);
    }
    __syncthreads();
    const int total_slots = s_slot_base[GP_NUM_CLASSES];
    const int pair_cta = tid / TPE;
    for (int it = 0; it < GP_BFS_CACHE_ITERS; ++it) {
        const int t0 = (gwarp + it * total_warps) * PPW;
        if (t0 >= total_slots) break;
        int cls = 0;
        while (t0 >= s_slot_base[cls + 1]) ++cls;
        const int gsh = cls <= 1 ? 4 : 5 - cls;
        const int rel = t0 - s_slot_base[cls] + pairlane;
        const int ent = s_ent_base[cls] + (rel >> gsh);
        const int sub = rel & ((1 << gsh) - 1);
        const bool active = ent < s_ent_base[cls + 1];
        const int4 d = active ? __ldg(p.desc + ent) : make_int4(0, 0, 0, -1);
        if (half == 0) {
            s_desc[it * PAIRS_CTA + pair_cta] = make_int4(d.x, d.z, d.w, sub | (gsh << 8) | ((int)active << 16));
            const int cnt = d.z & 0xFF;
#pragma unroll
            for (int i = 0; i < GP_SLOT_EDGES; ++i) {
                const int idx = sub + (i << gsh);
                // padding edges point back at the row itself (frontier[u] is a subset of seen[u])
                s_col[(it * GP_SLOT_EDGES + i) * PAIRS_CTA + pair_cta] = idx < cnt ? __ldg(p.col + d.y + idx) : d.x;
            }
            for (int b = 0; b < GP_BFS_DONE_BATCHES; ++b) s_done[(b * GP_BFS_CACHE_ITERS + it) * PAIRS_CTA + pair_cta] = 0;
        }
    }
    grid_barrier_any(p.bar + 0, bar_target[0], bar_prev_any[0], gridDim.x, false, &s_bcast, &s_any);

    int level = 1, max_level = 0;
    while (true) {
        LevelCtx<VW> c;
        // frontier of hop l lives in R[l] for l <= 15, then in the ping-pong pair
        auto frontier = [&](int l) -> u64 * {
            if (l == 0) return p.seeds;
            if (l <= GP_BFS_LEVEL_ARRAYS) return p.result + (size_t)l * p.plane_stride;
            return (l & 1) ? p.fr_a : p.fr_b;
        };
        c.cur = frontier(level - 1);
        c.nxt = frontier(level);
        c.seen = p.result;
        c.planes = p.result + (size_t)(1 + GP_BFS_LEVEL_ARRAYS) * p.plane_stride;
        c.plane_stride = p.plane_stride;
        c.level = level;
        // planes are first needed at hop 16 = 2^4: clear planes 0..4 one level ahead, and every
        // higher plane q one level before hop 2^q first sets it
        c.zero_first = 0;
        c.zero_count = 0;
        if (level == GP_BFS_LEVEL_ARRAYS) {
            c.zero_count = 5;
        } else if (level > GP_BFS_LEVEL_ARRAYS && ((level + 1) & level) == 0 && 31 - __clz(level + 1) < GP_BFS_PLANES) {
            c.zero_first = 31 - __clz(level + 1);
            c.zero_count = 1;
        }
        const u64 *live_r = p.live + (level % 3) * GP_BFS_MAX_LANE_WORDS;
        u64 *live_w = p.live + ((level + 1) % 3) * GP_BFS_MAX_LANE_WORDS;
        u64 *live_z = p.live + ((level + 2) % 3) * GP_BFS_MAX_LANE_WORDS;
        if (blockIdx.x == 0 && tid < lw) live_z[tid] = 0;
        GP_TRACE(0);

        for (int b = 0; b < p.batches; ++b) {
            u64 lv[VW], live_acc[VW];
#pragma unroll
            for (int i = 0; i < VW; ++i) {
                lv[i] = live_r[b * WB + woff + i];
                live_acc[i] = 0;
            }
            const u64 *cur_b = c.cur + (size_t)b * n * WB + woff;

            // ---- cached warp-iterations: descriptors / columns from shared memory, one round trip
            int it = 0;
            const int cached_iters = b < GP_BFS_DONE_BATCHES ? GP_BFS_CACHE_ITERS : 0;  // untracked batches stream
            for (; it < cached_iters; ++it) {
                const int tc = (gwarp + it * total_warps) * PPW;
                if (tc >= total_slots) break;
                const int4 cd = s_desc[it * PAIRS_CTA + pair_cta];
                const int sub = cd.w & 0xFF, gsh = (cd.w >> 8) & 0xFF;
                const bool active = (cd.w >> 16) & 1;
                const int nch = (cd.y >> 8) & 0x3FFFFF;
                const bool first_chunk = (cd.y >> 30) & 1;
                const bool track = b < GP_BFS_DONE_BATCHES;
                unsigned char *dflag = s_done + (b * GP_BFS_CACHE_ITERS + it) * PAIRS_CTA;
                // a row is "done" once every still-live lane has reached it: it never needs gathering again
                const bool done = track && dflag[pair_cta - sub] != 0;
                const bool work = active && !done;
                const size_t off = ((size_t)b * n + (size_t)cd.x) * WB + woff;
                u64 seenv[VW], acc[VW];
#pragma unroll
                for (int i = 0; i < VW; ++i) {
                    seenv[i] = ~0ull;
                    acc[i] = 0;
                }
                if (work) {
                    int v[GP_SLOT_EDGES];
#pragma unroll
                    for (int i = 0; i < GP_SLOT_EDGES; ++i) v[i] = s_col[(it * GP_SLOT_EDGES + i) * PAIRS_CTA + pair_cta];
                    u64 t[GP_SLOT_EDGES][VW];
#pragma unroll
                    for (int i = 0; i < GP_SLOT_EDGES; ++i) vload<VW>(cur_b + (size_t)(u32)v[i] * WB, t[i]);
                    if (sub == 0) vload<VW>(p.result + off, seenv);  // only the finalising lanes need it
#pragma unroll
                    for (int i = 0; i < GP_SLOT_EDGES; ++i)
#pragma unroll
                        for (int q = 0; q < VW; ++q) acc[q] |= t[i][q];
                    gathers += (u64)VW * GP_SLOT_EDGES;
                }
                for (int m = TPE; m < (TPE << gsh); m <<= 1)
#pragma unroll
                    for (int q = 0; q < VW; ++q) acc[q] |= shfl_xor_u64(acc[q], m);

                const bool leader = active && sub == 0;
                const u32 leader_mask = __ballot_sync(FULL_MASK, leader);
                if (leader) {
                    bool need = false;
#pragma unroll
                    for (int i = 0; i < VW; ++i) need |= work && (~seenv[i] & lv[i]) != 0;
                    bool need_row = need;
                    if constexpr (TPE == 2) need_row |= __shfl_xor_sync(leader_mask, (int)need, 1) != 0;
                    if (!need_row) {
                        // nothing left to reach here (monotone: seen grows, live shrinks): remember it
                        if (track && half == 0) dflag[pair_cta] = 1;
                        if (nch == 0 || first_chunk) {
#pragma unroll
                            for (int i = 0; i < VW; ++i) acc[i] = 0;
                            finalize_row<VW>(c, off, acc, seenv, live_acc);
                        }
                    } else if (nch == 0) {
                        finalize_row<VW>(c, off, acc, seenv, live_acc);
                    } else {
                        const size_t hidx = (size_t)b * p.hub_capacity + (size_t)cd.z;
                        u64 *accp = p.hub_acc + hidx * WB + woff;
                        u64 dep = 0;
#pragma unroll
                        for (int q = 0; q < VW; ++q)
                            if (acc[q]) dep |= atomicOr(accp + q, acc[q]);
                        u32 old = 0;
                        if constexpr (TPE == 2) {
                            dep |= __shfl_xor_sync(leader_mask, (u32)dep | (u32)(dep >> 32), 1);
                            asm volatile("" ::"l"(dep) : "memory");  // the ORs have returned from L2 before we count
                            if (half == 0) old = atomicAdd(p.hub_cnt + hidx, 1u);
                            old = __shfl_sync(leader_mask, old, lane & ~1);
                        } else {
                            asm volatile("" ::"l"(dep) : "memory");
                            old = atomicAdd(p.hub_cnt + hidx, 1u);
                        }
                        if (old == (u32)nch - 1u) {
                            u64 comb[VW];
#pragma unroll
                            for (int q = 0; q < VW; ++q) comb[q] = atomicExch(accp + q, 0ull);
                            if (half == 0) p.hub_cnt[hidx] = 0;
                            finalize_row<VW>(c, off, comb, seenv, live_acc);
                        }
                    }
                }
            }

            // ---- remaining warp-iterations (large graphs): streamed from global memory.
            // software pipeline: the descriptor of the next warp-iteration is fetched one iteration ahead
            int t0 = (gwarp + it * total_warps) * PPW;
            int cls = 0;
            int4 d_next = make_int4(0, 0, 0, -1);
            bool act_next = false;
            int sub_next = 0, gsh_next = 0;
            auto fetch = [&](int t) {
                while (t >= s_slot_base[cls + 1]) ++cls;  // regions are GP_SLOT_ALIGN aligned: warp-uniform
                gsh_next = cls <= 1 ? 4 : 5 - cls;        // log2 of slots per row: 16,16,8,4,2,1
                const int rel = t - s_slot_base[cls] + pairlane;
                const int ent = s_ent_base[cls] + (rel >> gsh_next);
                sub_next = rel & ((1 << gsh_next) - 1);
                act_next = ent < s_ent_base[cls + 1];
                d_next = act_next ? __ldg(p.desc + ent) : make_int4(0, 0, 0, -1);
            };
            if (t0 < total_slots) fetch(t0);
            while (t0 < total_slots) {
                const int4 d = d_next;
                const bool active = act_next;
                const int sub = sub_next, gsh = gsh_next;
                const int cnt = d.z & 0xFF, nch = (d.z >> 8) & 0x3FFFFF;
                const bool first_chunk = (d.z >> 30) & 1;
                const size_t off = ((size_t)b * n + (size_t)d.x) * WB + woff;
                // phase 1: this slot's column indices (sub, sub + G, ...: consecutive slots read consecutive
                // columns) and the row's seen words, all independent
                int v[GP_SLOT_EDGES];
#pragma unroll
                for (int i = 0; i < GP_SLOT_EDGES; ++i) {
                    // padding edges point back at the row itself: frontier[u] is a subset of seen[u],
                    // so it contributes nothing to acc & ~seen and the gathers below need no predicates
                    const int idx = sub + (i << gsh);
                    v[i] = idx < cnt ? __ldg(p.col + d.y + idx) : d.x;
                }
                u64 seenv[VW], acc[VW];
#pragma unroll
                for (int i = 0; i < VW; ++i) {
                    seenv[i] = ~0ull;
                    acc[i] = 0;
                }
                if (active) vload<VW>(p.result + off, seenv);
                t0 += total_warps * PPW;
                if (t0 < total_slots) fetch(t0);
                bool need = false;
#pragma unroll
                for (int i = 0; i < VW; ++i) need |= (~seenv[i] & lv[i]) != 0;
                // row-level verdict (both halves of the pair agree on it; used by the hub protocol)
                bool need_row = need;
                if constexpr (TPE == 2) need_row |= __shfl_xor_sync(FULL_MASK, (int)need, 1) != 0;
                // phase 2: all neighbour rows of the slot in flight at once
                if (need) {
                    u64 t[GP_SLOT_EDGES][VW];
#pragma unroll
                    for (int i = 0; i < GP_SLOT_EDGES; ++i) vload<VW>(cur_b + (size_t)(u32)v[i] * WB, t[i]);
#pragma unroll
                    for (int i = 0; i < GP_SLOT_EDGES; ++i)
#pragma unroll
                        for (int q = 0; q < VW; ++q) acc[q] |= t[i][q];
                    gathers += (u64)VW * (u64)max(0, min(GP_SLOT_EDGES, (cnt - sub + (1 << gsh) - 1) >> gsh));
                }
                // OR over the G slots of the row (warp-uniform trip count)
                for (int m = TPE; m < (TPE << gsh); m <<= 1)
#pragma unroll
                    for (int q = 0; q < VW; ++q) acc[q] |= shfl_xor_u64(acc[q], m);

                if (active && sub == 0) {
                    if (nch == 0) {
                        finalize_row<VW>(c, off, acc, seenv, live_acc);
                    } else if (!need_row) {
                        // hub row with nothing left to reach (every chunk sees the same seen / live
                        // words, so all agree): its first chunk alone writes the empty frontier
                        if (first_chunk) finalize_row<VW>(c, off, acc, seenv, live_acc);
                    } else {
                        // hub chunk: deposit the partial OR; the last chunk to arrive finalises the row.
                        // All traffic on hub_acc / hub_cnt is L2 atomics; waiting for the OR's return
                        // value orders it before the arrival count without a fence.
                        const size_t hidx = (size_t)b * p.hub_capacity + (size_t)d.w;
                        u64 *accp = p.hub_acc + hidx * WB + woff;
                        u64 dep = 0;
#pragma unroll
                        for (int q = 0; q < VW; ++q)
                            if (acc[q]) dep |= atomicOr(accp + q, acc[q]);
                        u32 old = 0;
                        if constexpr (TPE == 2) {
                            dep |= __shfl_xor_sync(LEADER_MASK, (u32)dep | (u32)(dep >> 32), 1);  // other half's returns
                            asm volatile("" ::"l"(dep) : "memory");  // the ORs have returned from L2 before we count
                            if (half == 0) old = atomicAdd(p.hub_cnt + hidx, 1u);
                            old = __shfl_sync(LEADER_MASK, old, 0);
                        } else {
                            asm volatile("" ::"l"(dep) : "memory");
                            old = atomicAdd(p.hub_cnt + hidx, 1u);
                        }
                        if (old == (u32)nch - 1u) {
                            u64 comb[VW];
#pragma unroll
                            for (int q = 0; q < VW; ++q) comb[q] = atomicExch(accp + q, 0ull);
                            if (half == 0) p.hub_cnt[hidx] = 0;
                            finalize_row<VW>(c, off, comb, seenv, live_acc);
                        }
                    }
                }
            }

            // ---- fold this batch's newly reached lanes into the CTA's live words
#pragma unroll
            for (int i = 0; i < VW; ++i) {
                u64 x = live_acc[i];
#pragma unroll
                for (int m = TPE; m < 32; m <<= 1) x |= shfl_xor_u64(x, m);
                if (lane < TPE && x) {
                    atomicOr(&s_live32[(b * WB + woff + i) * 2], (u32)x);
                    atomicOr(&s_live32[(b * WB + woff + i) * 2 + 1], (u32)(x >> 32));
                    s_any = 1;
                }
            }
        }
        GP_TRACE(3);
        __syncthreads();
        if (tid < lw) {
            const u64 x = ((u64)s_live32[tid * 2 + 1] << 32) | s_live32[tid * 2];
            if (x) atomicOr(live_w + tid, x);
            s_live32[tid * 2] = 0;
            s_live32[tid * 2 + 1] = 0;
        }
        const bool cta_any = s_any != 0;
        const int par = level & 1;
        const bool any = grid_barrier_any(p.bar + par, bar_target[par], bar_prev_any[par], gridDim.x, cta_any,
                                          &s_bcast, &s_any);
        if (!any) break;
        if (level >= (int)GP_UNREACHABLE_U16) {
            // a lane was first reached at hop 65535: not representable next to the 0xFFFF sentinel
            if (gtid == 0) atomicOr(&p.status[GP_BFS_ST_ERROR], GP_DEV_ERR_LEVEL_OVERFLOW);
            break;
        }
        max_level = level;
        ++level;
    }

    // ---- stats
    for (int m = 16; m; m >>= 1) gathers += shfl_xor_u64(gathers, m);
    if (lane == 0 && gathers) atomicAdd(p.counters, gathers);
    if (gtid == 0) {
        p.status[GP_BFS_ST_MAX_LEVEL] = max_level;
        p.status[GP_BFS_ST_LEVELS] = level;
        p.status[GP_BFS_ST_PULL] = level;
    }
}

template <int WB, int NT>
constexpr size_t bfs_cache_bytes()
{
    constexpr int pairs = NT / ((WB == 4) ? 2 : 1);
    return (size_t)GP_BFS_CACHE_ITERS * pairs * (sizeof(int4) + GP_SLOT_EDGES * sizeof(int)) +
           (size_t)GP_BFS_DONE_BATCHES * GP_BFS_CACHE_ITERS * pairs;
}

template <int WB, int NT, int MINB>
int launch_bfs_cfg(gp_msbfs *h, const BfsParams &p, cudaStream_t stream, int cfg_id)
{
    if (h->grid_blocks == 0 || h->grid_cfg != cfg_id * 8 + WB) {
        int occ = 0;
        GP_CUDA_CHECK(cudaFuncSetAttribute(msbfs_kernel<WB, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)bfs_cache_bytes<WB, NT>()));
        GP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, msbfs_kernel<WB, NT, MINB>, NT,
                                                                    bfs_cache_bytes<WB, NT>()));
        GP_REQUIRE(occ >= 1, GP_ERR_CUDA, "msbfs kernel does not fit on an SM");
        if (occ > MINB) occ = MINB;
        h->grid_blocks = occ * gp_sm_count();
        h->grid_cfg = cfg_id * 8 + WB;
        h->block_threads = NT;
    }
    BfsParams pp = p;
    void *args[] = {&pp};
    gp_count_launch();
    // inside a graph capture the timing events become event-record nodes (re-recorded on every replay)
    const unsigned ev_flags = gp_is_capturing() ? cudaEventRecordExternal : cudaEventRecordDefault;
    GP_CUDA_CHECK(cudaEventRecordWithFlags(h->ev_start, stream, ev_flags));
    GP_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)msbfs_kernel<WB, NT, MINB>, dim3(h->grid_blocks),
                                              dim3(NT), args, bfs_cache_bytes<WB, NT>(), stream));
    GP_CUDA_CHECK(cudaEventRecordWithFlags(h->ev_stop, stream, ev_flags));
    return GP_OK;
}

// Launch shapes (threads per CTA, CTAs per SM) trade resident warps against registers per thread,
// i.e. against how many neighbour gathers one thread keeps in flight.  GP_BFS_CFG picks one.
template <int WB>
int launch_bfs(gp_msbfs *h, const BfsParams &p, cudaStream_t stream)
{
    static int cfg = -1;
    if (cfg < 0) {
        const char *s = getenv("GP_BFS_CFG");
        cfg = s ? atoi(s) : GP_BFS_DEFAULT_CFG;
        if (cfg < 0 || cfg > 4) cfg = GP_BFS_DEFAULT_CFG;
    }
    switch (cfg) {
        case 0: return launch_bfs_cfg<WB, 512, 2>(h, p, stream, 0);   // 64 regs, 32 warps/SM
        case 1: return launch_bfs_cfg<WB, 384, 2>(h, p, stream, 1);   // 85 regs, 24 warps/SM
        case 2: return launch_bfs_cfg<WB, 512, 1>(h, p, stream, 2);   // 128 regs, 16 warps/SM
        case 3: return launch_bfs_cfg<WB, 1024, 1>(h, p, stream, 3);  // 64 regs, 32 warps/SM, half the CTAs
        default: return launch_bfs_cfg<WB, 256, 3>(h, p, stream, 4);  // 85 regs, 24 warps/SM
    }
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
    // result block R[0..31] (reached mask, 15 first-reached-at-hop arrays, 16 deep-hop bit planes),
    // then the seed frontier and the ping-pong pair
    alloc((void **)&h->seen, (size_t)GP_BFS_RESULT_ARRAYS * words * sizeof(u64));
    alloc((void **)&h->fr_a, 3 * words * sizeof(u64));
    alloc((void **)&h->live, 3 * GP_BFS_MAX_LANE_WORDS * sizeof(u64));
    h->hub_capacity = csr->hub_capacity;
    alloc((void **)&h->hub_acc, (size_t)h->hub_capacity * (size_t)h->cap_words_per_node * sizeof(u64));
    alloc((void **)&h->hub_cnt, (size_t)h->hub_capacity * (size_t)h->cap_words_per_node * sizeof(u32));
    alloc((void **)&h->bar, 8 * sizeof(u64));
    alloc((void **)&h->status, GP_BFS_ST_WORDS * sizeof(int));
    alloc((void **)&h->counters, 4 * sizeof(u64));
    if (getenv("GP_BFS_TRACE")) alloc((void **)&h->trace, GP_BFS_TRACE_WORDS * sizeof(u64));
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
    gp_drop_graphs(h);
    cudaFree(h->seen);
    cudaFree(h->fr_a);
    cudaFree(h->live);
    cudaFree(h->hub_acc);
    cudaFree(h->hub_cnt);
    cudaFree(h->bar);
    cudaFree(h->packed);
    cudaFree(h->deep_flag);
    cudaFree(h->status);
    cudaFree(h->counters);
    cudaFree(h->trace);
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
    h->seeds = h->fr_a + 2 * words;
    GP_CUDA_CHECK(cudaMemsetAsync(h->status, 0, GP_BFS_ST_WORDS * sizeof(int), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(h->counters, 0, 4 * sizeof(u64), stream));
    if (n == 0 || num_anchors == 0) {
        h->ran = true;
        return GP_OK;
    }
    GP_CUDA_CHECK(cudaMemsetAsync(h->seen, 0, words * sizeof(u64), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(h->seeds, 0, words * sizeof(u64), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(h->live, 0, 3 * GP_BFS_MAX_LANE_WORDS * sizeof(u64), stream));
    GP_CUDA_CHECK(cudaMemsetAsync(h->bar, 0, 8 * sizeof(u64), stream));
    if (!h->hub_zeroed) {
        // the kernel leaves these zeroed again (the finalising chunk resets its row's words)
        GP_CUDA_CHECK(cudaMemsetAsync(h->hub_acc, 0, (size_t)h->hub_capacity * h->cap_words_per_node * sizeof(u64), stream));
        GP_CUDA_CHECK(cudaMemsetAsync(h->hub_cnt, 0, (size_t)h->hub_capacity * h->cap_words_per_node * sizeof(u32), stream));
        h->hub_zeroed = true;
    }
    BfsParams p;
    p.n = (int)n;
    p.batches = batches;
    p.num_anchors = (int)num_anchors;
    p.hub_capacity = (int)h->hub_capacity;
    p.desc = h->csr->desc;
    p.hub_acc = h->hub_acc;
    p.hub_cnt = h->hub_cnt;
    p.bar = h->bar;
    p.col = h->csr->col_out;
    p.meta = h->csr->meta;
    p.anchors = (const long long *)d_anchors;
    p.result = h->seen;
    p.seeds = h->seeds;
    p.fr_a = h->fr_a;
    p.fr_b = h->fr_b;
    p.plane_stride = (long long)words;
    p.live = h->live;
    p.status = h->status;
    p.counters = h->counters;
    p.trace = h->trace;
    GP_TRY(wb == 1 ? launch_bfs<1>(h, p, stream) : wb == 2 ? launch_bfs<2>(h, p, stream)
                                                             : launch_bfs<4>(h, p, stream));
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

extern "C" int gp_msbfs_trace(gp_msbfs_t *h, uint64_t *h_out, int64_t cap_words, int32_t *levels, int32_t *warps)
{
    GP_REQUIRE(h != nullptr && h_out != nullptr && levels && warps, GP_ERR_INVALID, "gp_msbfs_trace: NULL argument");
    GP_REQUIRE(h->trace != nullptr, GP_ERR_INVALID, "gp_msbfs_trace: set GP_BFS_TRACE=1 before gp_msbfs_create");
    gp_msbfs_stats_t st;
    gp_msbfs_stats(h, &st, nullptr);
    GP_CUDA_CHECK(cudaDeviceSynchronize());
    *levels = st.levels_run < 32 ? st.levels_run : 32;
    *warps = h->grid_blocks * (h->block_threads / 32);
    const int64_t words = (int64_t)(*levels) * (*warps) * 4;
    GP_REQUIRE(words <= cap_words && words <= GP_BFS_TRACE_WORDS, GP_ERR_INVALID, "gp_msbfs_trace: buffer too small");
    GP_CUDA_CHECK(cudaMemcpy(h_out, h->trace, sizeof(u64) * (size_t)words, cudaMemcpyDeviceToHost));
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
    *d_planes = (const uint64_t *)h->seen;
    *plane_stride_words = (int64_t)h->wb * h->batches * h->num_nodes;
    *num_planes = st.max_level <= GP_BFS_LEVEL_ARRAYS ? 1 + st.max_level : GP_BFS_RESULT_ARRAYS;
    *batches = h->batches;
    *words_per_batch = h->wb;
    return GP_OK;
}
