// gp_internal.h — handle layouts shared by the .cu translation units (not part of the ABI).
#pragma once

#include "gp_common.cuh"

// Work list of the MS-BFS (see gp_msbfs.cu).  A "slot" is one thread gathering at most
// GP_SLOT_EDGES neighbour rows; a row of degree d is served by G = 1,2,4,8,16,32 slots
// (G * GP_SLOT_EDGES >= d), rows above 32 * GP_SLOT_EDGES edges are cut into chunks of that
// size (class 0), one warp-wide tile each.
constexpr int GP_SLOT_EDGES = 4;
constexpr int GP_CHUNK_EDGES = 32 * GP_SLOT_EDGES;  // 128
constexpr int GP_NUM_CLASSES = 7;   // 0: chunks of hub rows (G=32), 1: G=32, 2: G=16, 3: G=8, 4: G=4, 5: G=2, 6: G=1
constexpr int GP_SLOT_ALIGN = 32;   // class regions start on a multiple of this many slots (one warp tile)

static_assert(GP_CHUNK_EDGES == 128, "degree_class() in gp_csr.cu hard-codes the class thresholds");

// Device-resident metadata words of a CSR (int32 each).
enum : int {
    GP_META_NUM_EDGES = 0,    // E' after de-duplication
    GP_META_ERROR = 1,        // GP_DEV_ERR_* bits
    GP_META_NUM_BIG_ROWS = 2, // rows queued for the CTA-wide row sort (more than 128 raw edges)
    GP_META_SCRATCH = 3,
    GP_META_MAX_DEGREE = 4,
    GP_META_IS_SYMMETRIC = 5,
    GP_META_IN_BUILT = 6,
    GP_META_NUM_HUB_ROWS = 7,   // rows with degree > GP_CHUNK_EDGES
    GP_META_NUM_MED_ROWS = 8,   // rows of 17..128 raw edges queued for the warp-per-row sort (gp_csr_build)
    GP_META_ENT_BASE = 16,      // [8]: first descriptor of each class, then the total
    GP_META_SLOT_BASE = 24,     // [8]: first slot of each class, then the total
    GP_META_WORDS = 32
};

struct gp_csr {
    uint64_t uid = gp_next_uid();
    int64_t num_nodes = 0;
    int64_t edge_capacity = 0;  // input columns accepted by gp_csr_build
    int64_t key_capacity = 0;   // edge_capacity * (symmetrize ? 2 : 1)
    int64_t num_input_edges = 0;
    uint32_t flags = 0;
    int64_t hub_capacity = 0;   // upper bound on rows with degree > GP_CHUNK_EDGES
    int64_t desc_capacity = 0;  // work-list descriptors: the sum of the class regions below
    int desc_off[8] = {};       // first descriptor of each degree class's region (fixed at create: a class cannot
                                // hold more rows than edge_capacity / its smallest degree), then the total
    int64_t big_capacity = 0;   // upper bound on rows with more than 128 raw edges
    bool built = false;
    bool in_built = false;      // host view of GP_META_IN_BUILT (in-edge CSR materialised)

    // out-edge CSR, "gapped": row r = col[row_start[r] .. row_start[r] + deg[r]), ascending, unique
    int *deg = nullptr;         // [N + 1] raw edge count per row, then the distinct out-degree
    int *row_start = nullptr;   // [N + 1] first column of each row (exclusive prefix of the RAW counts)
    int *col = nullptr;         // [key_capacity]
    int *cursor = nullptr;      // [N + 1] scatter cursors
    // in-edge CSR (compact), allocated and built on first use
    int *rowptr_in = nullptr;   // [N + 1]
    int *col_in = nullptr;      // [key_capacity]
    int *deg_in = nullptr;      // [N + 1]
    int *biglist = nullptr;     // [big_capacity] rows queued for the CTA-wide row sort
    int *medlist = nullptr;     // [med_capacity] rows of 17..128 raw edges, sorted by one warp each beside the short rows
    int64_t med_capacity = 0;
    int4 *desc = nullptr;       // [desc_capacity] {row, first edge, count | chunks << 8, hub index or -1}
    int *meta = nullptr;        // [GP_META_WORDS], followed in the same allocation by scan_status
    int *scan_status = nullptr; // look-back words of the chained scans + ticket counters in the last 8 words
    size_t scan_status_words = 0;
    size_t scan_b_offset = 0;   // first status word of the second scan of a build
    // grow-only device scratch owned by the handle (gp_csr_scratch), freed by gp_csr_free: samplers that run for
    // seconds (betweenness) keep their workspace here instead of paying cudaMalloc / cudaFree on every call
    void *scratch[16] = {};
    size_t scratch_bytes[16] = {};
    cudaStream_t side = nullptr;    // the long rows are sorted beside the short ones (sort_rows_forked)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t trace_ev[12] = {};  // GP_CSR_TRACE: events after every launch of the last build (diagnostics)
    int trace_n = 0;
    int bitmap_words = 0;       // ceil(N / 32) if the long-row sort may use a node bitmap in shared memory, else 0
    int big_smem_bytes = 0;     // dynamic shared memory of rowsort_big_kernel
};

// Device scratch slot `slot` of at least `bytes` bytes (contents undefined; re-allocated only when it must grow).
int gp_csr_scratch(gp_csr *csr, int slot, size_t bytes, void **out);
// Makes sure the in-edge CSR exists (transpose sort unless symmetric).  Async.
int gp_csr_ensure_in(gp_csr *csr, cudaStream_t stream);
