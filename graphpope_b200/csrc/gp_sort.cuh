// gp_sort.cuh — device-wide LSD radix sort ("onesweep": one upfront histogram, then one
// chained-scan scatter kernel per digit) and a fused unique-compaction, both hand-written
// for sm_100a.  Element counts may live on the device (d_n) so a whole pipeline can be
// enqueued / graph-captured without a host round trip.
#pragma once

#include "gp_common.cuh"

constexpr int GP_SORT_MAX_PASSES = 8;
constexpr int GP_SORT_THREADS = 256;
constexpr int GP_SORT_IPT = 8;                              // keys per thread
constexpr int GP_SORT_TILE = GP_SORT_THREADS * GP_SORT_IPT; // keys per tile

struct GpSortWorkspace {
    u64 *keys_alt = nullptr;   // [capacity]
    u32 *vals_alt = nullptr;   // [capacity] (only if with_values)
    u32 *scratch = nullptr;    // hist [8][256] | tile counters [8] | status [passes][tiles][256]
    size_t scratch_bytes = 0;
    int64_t capacity = 0;
    int64_t max_tiles = 0;
    bool with_values = false;
};

int gp_sort_workspace_create(GpSortWorkspace *ws, int64_t capacity, bool with_values);
void gp_sort_workspace_free(GpSortWorkspace *ws);

// Sorts keys (and optional u32 values) ascending on key bits [bit_lo, bit_hi), stable.
// n_max bounds the grid; the live count is *d_n if d_n != nullptr, else n_max.
// The sorted data ends up in either the input buffers or the workspace's alternates;
// *keys_out / *vals_out tell which.  Async on `stream`.
int gp_radix_sort(GpSortWorkspace *ws, u64 *keys, u32 *vals, const u32 *d_n, int64_t n_max,
                  int bit_lo, int bit_hi, cudaStream_t stream, u64 **keys_out, u32 **vals_out);

// out[0..m) = the distinct keys of the sorted array in[0..n) in order; *d_m = m.
// `status` needs gp_ceil_div(n_max, GP_SORT_TILE) + 1 zeroed u32 words (the function zeroes them).
int gp_unique_sorted(const u64 *in, u64 *out, const u32 *d_n, int64_t n_max, u32 *d_m,
                     u32 *status, cudaStream_t stream);
