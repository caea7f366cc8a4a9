// gp_msbfs.cuh — handle layout of the multi-source BFS (internal).
#pragma once

#include "gp_internal.h"

constexpr int GP_BFS_THREADS = 512;
constexpr int GP_BFS_MAX_LANE_WORDS = 256;  // B * WB cap (K <= 16384 per GPU)
constexpr int GP_BFS_PLANES = 16;           // distance bit planes (uint16 range)

enum : int {
    GP_BFS_ST_MAX_LEVEL = 0,
    GP_BFS_ST_ERROR = 1,
    GP_BFS_ST_LEVELS = 2,
    GP_BFS_ST_PULL = 3,
    GP_BFS_ST_PUSH = 4,
    GP_BFS_ST_WORDS = 8
};

struct gp_msbfs {
    const gp_csr *csr = nullptr;
    int64_t num_nodes = 0;
    int64_t max_anchors = 0;
    int64_t cap_words_per_node = 0;  // lane words per node the buffers were sized for
    // configuration of the last run
    int64_t num_anchors = 0;
    int wb = 1;        // lane words per batch row (1, 2 or 4)
    int batches = 0;   // independent 64*wb-anchor batches
    bool ran = false;

    u64 *seen = nullptr;     // [batches][N][wb]   "reached" masks (plane 0 of the result)
    u64 *fr_a = nullptr;     // frontier ping
    u64 *fr_b = nullptr;     // frontier pong
    u64 *planes = nullptr;   // [GP_BFS_PLANES][batches][N][wb] bit-sliced hop distance
    u64 *live = nullptr;     // [3][GP_BFS_MAX_LANE_WORDS]
    int *queue = nullptr;    // [N] compacted active rows (push levels)
    u32 *sync_words = nullptr;  // [0] grid barrier, [1..] queue cursors / counters
    int *status = nullptr;      // [GP_BFS_ST_WORDS]
    u64 *counters = nullptr;    // [4] gathers issued, pushes issued, ...
    int grid_blocks = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;  // bracket the persistent kernel alone
};

// Fused decode + concat epilogue over `num_ranks` plane sets (1 = local result).
struct GpDecodeParams {
    const u64 *planes0;        // plane set of rank 0: [num_planes_alloc][batches][N][wb], plane 0 = reached
    long long rank_stride;     // words between consecutive ranks' plane sets
    long long plane_stride;    // words between consecutive planes
    int num_ranks;
    int num_dist_planes;       // distance bit planes to read (bit_length(max_level)); <0: read from status
    const int *status;         // device status words (max level) when num_dist_planes < 0
    long long n;
    long long anchors_per_rank;
    int wb;
    const float *x;            // optional [N, ld_x]
    long long num_features, ld_x;
    float *out;
    long long ld_out, col_offset;
};
int gp_launch_decode_features(const GpDecodeParams &p, cudaStream_t stream);
