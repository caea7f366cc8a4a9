// gp_internal.h — handle layouts shared by the .cu translation units (not part of the ABI).
#pragma once

#include "gp_common.cuh"
#include "gp_sort.cuh"

// Work list of the MS-BFS (see gp_msbfs.cu).  A "slot" is one thread gathering at most
// GP_SLOT_EDGES neighbour rows; a row of degree d is served by G = 1,2,4,8,16,32 slots
// (G * GP_SLOT_EDGES >= d), rows above 32 * GP_SLOT_EDGES edges are cut into chunks of that
// size (class 0), one warp-wide tile each.
constexpr int GP_SLOT_EDGES = 4;
constexpr int GP_CHUNK_EDGES = 32 * GP_SLOT_EDGES;  // 128
constexpr int GP_NUM_CLASSES = 7;   // 0: chunks of hub rows (G=32), 1: G=32, 2: G=16, 3: G=8, 4: G=4, 5: G=2, 6: G=1
constexpr int GP_SLOT_ALIGN = 32;   // class regions start on a multiple of this many slots (one warp tile)

// Device-resident metadata words of a CSR (int32 each).
enum : int {
    GP_META_NUM_EDGES = 0,    // E' after de-duplication
    GP_META_ERROR = 1,        // GP_DEV_ERR_* bits
    GP_META_MAX_DEGREE = 4,
    GP_META_IS_SYMMETRIC = 5,
    GP_META_IN_BUILT = 6,
    GP_META_NUM_HUB_ROWS = 7,   // rows with degree > GP_CHUNK_EDGES
    GP_META_RANK = 8,           // [7]: rows (in degree order) with degree > 128, 64, 32, 16, 8, 4, then N
    GP_META_ENT_BASE = 16,      // [8]: first descriptor of each class, then the total
    GP_META_SLOT_BASE = 24,     // [8]: first slot of each class, then the total
    GP_META_WORDS = 32
};

struct gp_csr {
    int64_t num_nodes = 0;
    int64_t edge_capacity = 0;  // input columns accepted by gp_csr_build
    int64_t key_capacity = 0;   // edge_capacity * (symmetrize ? 2 : 1)
    int64_t num_input_edges = 0;
    uint32_t flags = 0;
    int node_bits = 1;          // bits needed for a node id
    int64_t hub_capacity = 0;   // upper bound on rows with degree > GP_CHUNK_EDGES
    int64_t desc_capacity = 0;  // upper bound on work-list descriptors
    bool built = false;
    bool in_built = false;      // host view of GP_META_IN_BUILT (in-edge CSR materialised)

    u64 *keys = nullptr;        // [key_capacity] packed (src << node_bits | dst)
    u64 *ukeys = nullptr;       // [key_capacity] sorted unique keys
    int *rowptr_out = nullptr;  // [N + 1]
    int *col_out = nullptr;     // [key_capacity]
    int *rowptr_in = nullptr;   // [N + 1]  (== rowptr_out when the graph is symmetric)
    int *col_in = nullptr;      // [key_capacity]
    int *order = nullptr;       // [N] node ids by descending out-degree (ties: ascending id)
    int4 *desc = nullptr;       // [desc_capacity] {row, first edge, count | chunks << 8, hub index or -1}
    int *hub_chunk_off = nullptr;  // [hub_capacity + 1] first descriptor of each hub row
    u64 *okeys = nullptr;       // [N] sort keys for `order`
    int *meta = nullptr;        // [GP_META_WORDS]
    u32 *uniq_status = nullptr; // look-back words for gp_unique_sorted
    GpSortWorkspace sort_ws;
};

// Makes sure the in-edge CSR exists (transpose sort unless symmetric).  Async.
int gp_csr_ensure_in(gp_csr *csr, cudaStream_t stream);
