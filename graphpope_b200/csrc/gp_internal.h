// gp_internal.h — handle layouts shared by the .cu translation units (not part of the ABI).
#pragma once

#include "gp_common.cuh"
#include "gp_sort.cuh"

// Row classes of the degree-ordered work list (see gp_msbfs.cu).
constexpr int GP_DEG_SMALL_MAX_DEFAULT = 16;   // <= : one thread group per row
constexpr int GP_DEG_LARGE_MIN_DEFAULT = 512;  // >= : one CTA per row; between: one warp per row

// Device-resident metadata words of a CSR (int32 each).
enum : int {
    GP_META_NUM_EDGES = 0,    // E' after de-duplication
    GP_META_ERROR = 1,        // GP_DEV_ERR_* bits
    GP_META_N_LARGE = 2,      // rows with degree >= large_min          (order[0 .. n_large))
    GP_META_N_LARGE_MED = 3,  // rows with degree >  small_max          (order[0 .. n_large_med))
    GP_META_MAX_DEGREE = 4,
    GP_META_IS_SYMMETRIC = 5,
    GP_META_IN_BUILT = 6,
    GP_META_WORDS = 16
};

struct gp_csr {
    int64_t num_nodes = 0;
    int64_t edge_capacity = 0;  // input columns accepted by gp_csr_build
    int64_t key_capacity = 0;   // edge_capacity * (symmetrize ? 2 : 1)
    int64_t num_input_edges = 0;
    uint32_t flags = 0;
    int node_bits = 1;          // bits needed for a node id
    int deg_small_max = GP_DEG_SMALL_MAX_DEFAULT;
    int deg_large_min = GP_DEG_LARGE_MIN_DEFAULT;
    bool built = false;
    bool in_built = false;      // host view of GP_META_IN_BUILT (in-edge CSR materialised)

    u64 *keys = nullptr;        // [key_capacity] packed (src << node_bits | dst)
    u64 *ukeys = nullptr;       // [key_capacity] sorted unique keys
    int *rowptr_out = nullptr;  // [N + 1]
    int *col_out = nullptr;     // [key_capacity]
    int *rowptr_in = nullptr;   // [N + 1]  (== rowptr_out when the graph is symmetric)
    int *col_in = nullptr;      // [key_capacity]
    int *order = nullptr;       // [N] node ids by descending out-degree (ties: ascending id)
    u64 *okeys = nullptr;       // [N] sort keys for `order`
    int *meta = nullptr;        // [GP_META_WORDS]
    u32 *uniq_status = nullptr; // look-back words for gp_unique_sorted
    GpSortWorkspace sort_ws;
};

// Makes sure the in-edge CSR exists (transpose sort unless symmetric).  Async.
int gp_csr_ensure_in(gp_csr *csr, cudaStream_t stream);
