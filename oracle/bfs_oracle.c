/*
 * bfs_oracle.c — plain-C restatement of the geodesic GraphPOPE hot loop.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA
 * path and the "fair CPU" baseline of bench.py.  Never linked into the product.
 *
 * What it restates (citations into /root/reference):
 *   utils.py:121      to_networkx(data): DiGraph, parallel edges collapse,
 *                     self-loops kept, NO symmetrisation  -> gpo_build_in_csr
 *   utils.py:69-77    for node, for anchor: 1/len(shortest_path(node, anchor)),
 *                     0 when no path.  len(path) = hops + 1, and hops(node ->
 *                     anchor) over all nodes is one BFS from the anchor over
 *                     REVERSED edges (in-neighbour lists)  -> gpo_bfs_hops
 *   utils.py:73,76,125  value convention 1/(d+1) in float64 rounded to float32,
 *                     unreachable = 0                       -> gpo_normalise
 *
 * Algorithmic difference from the reference, stated plainly: the reference runs
 * N*K bidirectional BFS calls; this runs K single-source BFS runs.  The
 * outputs are equal (tests/test_oracle_geodesic.py pins T0 == T1 == T2 == C on
 * the golden fixtures produced by the reference itself).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GPO_UNREACHABLE 0xFFFFu

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

/*
 * In-edge CSR of the de-duplicated digraph: row v lists every u with an edge
 * u -> v, ascending, unique.  rowptr has n+1 entries, col has *e_unique.
 * Caller frees both with gpo_free.  Returns 0, or -1 on a bad index / OOM.
 */
int gpo_build_in_csr(const int64_t *edge_index, int64_t e, int64_t n, int symmetrize,
                     int64_t **rowptr_out, int32_t **col_out, int64_t *e_unique)
{
    int64_t m = symmetrize ? 2 * e : e;
    uint64_t *code = (uint64_t *)malloc((size_t)(m > 0 ? m : 1) * sizeof(uint64_t));
    int64_t *rowptr = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    if (!code || !rowptr) { free(code); free(rowptr); return -1; }
    const int64_t *src = edge_index, *dst = edge_index + e;
    for (int64_t i = 0; i < e; ++i) {
        int64_t u = src[i], v = dst[i];
        if (u < 0 || v < 0 || u >= n || v >= n) { free(code); free(rowptr); return -1; }
        code[i] = ((uint64_t)v << 32) | (uint64_t)u;           /* key (dst, src) */
        if (symmetrize) code[e + i] = ((uint64_t)u << 32) | (uint64_t)v;
    }
    qsort(code, (size_t)m, sizeof(uint64_t), cmp_u64);
    int64_t uq = 0;
    for (int64_t i = 0; i < m; ++i)
        if (i == 0 || code[i] != code[i - 1]) code[uq++] = code[i];
    int32_t *col = (int32_t *)malloc((size_t)(uq > 0 ? uq : 1) * sizeof(int32_t));
    if (!col) { free(code); free(rowptr); return -1; }
    for (int64_t i = 0; i < uq; ++i) {
        rowptr[(code[i] >> 32) + 1] += 1;
        col[i] = (int32_t)(code[i] & 0xFFFFFFFFu);
    }
    for (int64_t v = 0; v < n; ++v) rowptr[v + 1] += rowptr[v];
    free(code);
    *rowptr_out = rowptr; *col_out = col; *e_unique = uq;
    return 0;
}

void gpo_free(void *p) { free(p); }

/*
 * One queue BFS per anchor over the in-edge CSR.  dist is [n, ld] uint16,
 * column col_offset + j receives anchor j; unreachable = 0xFFFF.
 * Returns the largest finite hop count seen, -1 on OOM, -2 if a distance
 * would not fit uint16.
 */
int64_t gpo_bfs_hops(const int64_t *rowptr, const int32_t *col, int64_t n,
                     const int64_t *anchors, int64_t k,
                     uint16_t *dist, int64_t ld, int64_t col_offset)
{
    int32_t *queue = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    int32_t *d = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    if (!queue || !d) { free(queue); free(d); return -1; }
    int64_t max_level = 0;
    for (int64_t j = 0; j < k; ++j) {
        for (int64_t v = 0; v < n; ++v) d[v] = -1;
        int64_t a = anchors[j];
        if (a < 0 || a >= n) { free(queue); free(d); return -1; }
        int64_t head = 0, tail = 0;
        d[a] = 0; queue[tail++] = (int32_t)a;
        while (head < tail) {
            int32_t v = queue[head++];
            int32_t dv = d[v];
            for (int64_t p = rowptr[v]; p < rowptr[v + 1]; ++p) {
                int32_t u = col[p];
                if (d[u] < 0) { d[u] = dv + 1; queue[tail++] = u; }
            }
        }
        for (int64_t v = 0; v < n; ++v) {
            int32_t dv = d[v];
            if (dv >= (int32_t)GPO_UNREACHABLE) { free(queue); free(d); return -2; }
            if (dv > max_level) max_level = dv;
            dist[v * ld + col_offset + j] = dv < 0 ? (uint16_t)GPO_UNREACHABLE : (uint16_t)dv;
        }
    }
    free(queue); free(d);
    return max_level;
}

/* utils.py:73,76,125: float64 1/(d+1) rounded to float32; unreachable -> 0. */
void gpo_normalise(const uint16_t *dist, int64_t count, float *out)
{
    for (int64_t i = 0; i < count; ++i)
        out[i] = dist[i] == GPO_UNREACHABLE ? 0.0f : (float)(1.0 / ((double)dist[i] + 1.0));
}
