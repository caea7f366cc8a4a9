/* A plain-C caller of the drop-in boundary (include/graphpope_b200.h): no Python, no torch, host buffers only.
 *
 *   gcc -std=c99 -Iinclude examples/embed_host.c -Lgraphpope_b200 -lgraphpope_b200 \
 *       -Wl,-rpath,$PWD/graphpope_b200 -o /tmp/embed_host && /tmp/embed_host
 *
 * It runs gp_geodesic_embed_host — what the reference does in get_geodesic_distance_vector + concat_into_features
 * (utils.py:116-135) — on the directed graph below and checks the rows against the reference's conventions
 * (utils.py:64-81): value 1/(hops+1), 1.0 for a node that is its own anchor, 0.0 when the anchor cannot be reached,
 * edges followed in their own direction only, duplicate edges and self loops harmless, anchors may repeat.
 *
 *   0 -> 1 -> 2 -> 3      4 -> 2      5 (isolated)      1 -> 1 (self loop)      0 -> 1 twice
 *
 * Exit status: 0 = all rows as expected, 1 = a mismatch, 2 = the library reported an error (for instance
 * GP_ERR_NO_DEVICE on a machine without a B200; the message is printed).                                        */
#include <stdio.h>
#include <string.h>

#include "graphpope_b200.h"

#define N 6
#define E 7
#define K 4
#define F 2

int main(void) {
    /* edge_index [2, E], row 0 = source, row 1 = target, as torch_geometric stores it */
    const int64_t edge_index[2 * E] = {0, 1, 2, 4, 1, 0, 0,
                                       1, 2, 3, 2, 1, 1, 1};
    const int64_t anchors[K] = {3, 0, 2, 3}; /* the first and the last column must come out identical */
    float x[N * F];
    float out[N * (F + K)];
    uint16_t hops[N * K];
    /* expected hop counts, 65535 = unreachable (row = source node, column = anchor) */
    const uint16_t want_hops[N * K] = {3, 0, 2, 3,             /* node 0 */
                                       2, 65535, 1, 2,         /* node 1: nothing leads back to 0 */
                                       1, 65535, 0, 1,         /* node 2 */
                                       0, 65535, 65535, 0,     /* node 3: a sink */
                                       2, 65535, 1, 2,         /* node 4 */
                                       65535, 65535, 65535, 65535}; /* node 5 */
    gp_msbfs_stats_t stats;
    int i, j, rc, bad = 0;

    for (i = 0; i < N * F; ++i) x[i] = (float)i + 0.5f;
    memset(out, 0xff, sizeof out);
    memset(&stats, 0, sizeof stats);

    printf("graphpope_b200 ABI %d\n", gp_abi_version());
    rc = gp_geodesic_embed_host(edge_index, E, N, 0u, anchors, K, x, F, out, F + K, F, hops, &stats);
    if (rc != GP_OK) {
        printf("gp_geodesic_embed_host: status %d (%s): %s\n", rc, gp_status_string(rc), gp_last_error());
        return 2;
    }
    for (i = 0; i < N; ++i) {
        for (j = 0; j < F; ++j)
            if (out[i * (F + K) + j] != x[i * F + j]) ++bad;
        for (j = 0; j < K; ++j) {
            const uint16_t h = want_hops[i * K + j];
            const float want = h == 65535 ? 0.0f : 1.0f / (float)(h + 1);
            if (hops[i * K + j] != h || out[i * (F + K) + F + j] != want) {
                printf("node %d anchor %d: hops %u (want %u), value %.9g (want %.9g)\n", i, (int)anchors[j],
                       (unsigned)hops[i * K + j], (unsigned)h, (double)out[i * (F + K) + F + j], (double)want);
                ++bad;
            }
        }
    }
    printf("%d nodes x (%d features + %d anchors): %s\n", N, F, K, bad ? "MISMATCH" : "ok");
    return bad ? 1 : 0;
}
