// gather_rate.cu — how fast can one B200 gather random 32-byte rows out of an L2-resident table?
// This is the inner operation of a dense MS-BFS level (one row = 256 anchor lanes of a neighbour).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_rate gather_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int NT = 384;
constexpr int BATCH = 8;

// V0: lane pair per row, LDG.128 each (16 rows per warp-wide load)
template <bool CG>
__global__ void __launch_bounds__(NT, 2) v_pair128(const uint4 *__restrict__ tab, const int *__restrict__ idx, long long m, u64 *out)
{
    const int lane = threadIdx.x & 31, half = lane & 1, pl = lane >> 1;
    const long long gw = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5), tw = (long long)gridDim.x * (NT / 32);
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (long long base = gw * 16 * BATCH; base + 16 * BATCH <= m; base += tw * 16 * BATCH) {
        int v[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) v[i] = idx[base + i * 16 + pl];
        uint4 t[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const uint4 *p = tab + (size_t)v[i] * 2 + half;
            if (CG) t[i] = __ldcg(p); else t[i] = *p;
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i) { acc.x |= t[i].x; acc.y |= t[i].y; acc.z |= t[i].z; acc.w |= t[i].w; }
    }
    if ((acc.x | acc.y | acc.z | acc.w) == 0x12345678u) out[0] = acc.x;
}

// V1: LPR lanes per row, each loads 32/LPR bytes (LPR = 4 -> LDG.64, 8 -> LDG.32)
template <int LPR>
__global__ void __launch_bounds__(NT, 2) v_lanes(const u32 *__restrict__ tab, const int *__restrict__ idx, long long m, u64 *out)
{
    constexpr int RPW = 32 / LPR;  // rows per warp-wide load
    constexpr int WORDS = 8 / LPR; // u32 per lane
    const int lane = threadIdx.x & 31, sub = lane % LPR, pl = lane / LPR;
    const long long gw = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5), tw = (long long)gridDim.x * (NT / 32);
    u32 acc[WORDS] = {0};
    for (long long base = gw * RPW * BATCH; base + RPW * BATCH <= m; base += tw * RPW * BATCH) {
        int v[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) v[i] = idx[base + i * RPW + pl];
        u32 t[BATCH][WORDS];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const u32 *p = tab + (size_t)v[i] * 8 + sub * WORDS;
            if (WORDS == 2) { uint2 x = *reinterpret_cast<const uint2 *>(p); t[i][0] = x.x; t[i][WORDS - 1] = x.y; }
            else t[i][0] = *p;
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i)
#pragma unroll
            for (int q = 0; q < WORDS; ++q) acc[q] |= t[i][q];
    }
    u32 a = 0;
    for (int q = 0; q < WORDS; ++q) a |= acc[q];
    if (a == 0x12345678u) out[0] = a;
}

// V3: thread per row, two LDG.128 (or one 256-bit load)
template <bool WIDE>
__global__ void __launch_bounds__(NT, 2) v_thread(const uint4 *__restrict__ tab, const int *__restrict__ idx, long long m, u64 *out)
{
    constexpr int B2 = BATCH / 2;
    const long long gt = (long long)blockIdx.x * NT + threadIdx.x, tt = (long long)gridDim.x * NT;
    const int lane = threadIdx.x & 31;
    const long long gw = gt >> 5, tw = tt >> 5;
    u64 acc[4] = {0, 0, 0, 0};
    for (long long base = gw * 32 * B2; base + 32 * B2 <= m; base += tw * 32 * B2) {
        int v[B2];
#pragma unroll
        for (int i = 0; i < B2; ++i) v[i] = idx[base + i * 32 + lane];
        u64 t[B2][4];
#pragma unroll
        for (int i = 0; i < B2; ++i) {
            const uint4 *p = tab + (size_t)v[i] * 2;
            if (WIDE) {
                asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(t[i][0]), "=l"(t[i][1]), "=l"(t[i][2]), "=l"(t[i][3]) : "l"(p));
            } else {
                const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(p), b = *reinterpret_cast<const ulonglong2 *>(p + 1);
                t[i][0] = a.x; t[i][1] = a.y; t[i][2] = b.x; t[i][3] = b.y;
            }
        }
#pragma unroll
        for (int i = 0; i < B2; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] |= t[i][q];
    }
    if ((acc[0] | acc[1] | acc[2] | acc[3]) == 0x12345678ull) out[0] = acc[0];
}

// V6: one LDG per row with only ONE lane pair active per instruction (16 single-line LDGs instead of one 16-line LDG)
__global__ void __launch_bounds__(NT, 2) v_single(const uint4 *__restrict__ tab, const int *__restrict__ idx, long long m, u64 *out)
{
    const int lane = threadIdx.x & 31, half = lane & 1, pl = lane >> 1;
    const long long gw = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5), tw = (long long)gridDim.x * (NT / 32);
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (long long base = gw * 16 * BATCH; base + 16 * BATCH <= m; base += tw * 16 * BATCH) {
        int v[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) v[i] = idx[base + i * 16 + pl];
        uint4 t[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            t[i] = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int s = 0; s < 16; ++s)
                if (pl == s) t[i] = tab[(size_t)v[i] * 2 + half];
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i) { acc.x |= t[i].x; acc.y |= t[i].y; acc.z |= t[i].z; acc.w |= t[i].w; }
    }
    if ((acc.x | acc.y | acc.z | acc.w) == 0x12345678u) out[0] = acc.x;
}

// V7: cp.async (LDGSTS) 16 B per lane, pair per row, into shared memory; then read back
__global__ void __launch_bounds__(NT, 2) v_ldgsts(const uint4 *__restrict__ tab, const int *__restrict__ idx, long long m, u64 *out)
{
    __shared__ uint4 buf[BATCH][NT];
    const int lane = threadIdx.x & 31, half = lane & 1, pl = lane >> 1;
    const long long gw = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5), tw = (long long)gridDim.x * (NT / 32);
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (long long base = gw * 16 * BATCH; base + 16 * BATCH <= m; base += tw * 16 * BATCH) {
        int v[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) v[i] = idx[base + i * 16 + pl];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const u32 dst = (u32)__cvta_generic_to_shared(&buf[i][threadIdx.x]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(tab + (size_t)v[i] * 2 + half) : "memory");
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int i = 0; i < BATCH; ++i) { const uint4 t = buf[i][threadIdx.x]; acc.x |= t.x; acc.y |= t.y; acc.z |= t.z; acc.w |= t.w; }
    }
    if ((acc.x | acc.y | acc.z | acc.w) == 0x12345678u) out[0] = acc.x;
}

// V8: TMA bulk copy, 32 B per row, one thread per row, mbarrier completion
__global__ void __launch_bounds__(NT, 2) v_bulk(const uint4 *__restrict__ tab, const int *__restrict__ idx, long long m, u64 *out)
{
    constexpr int B2 = 3;
    __shared__ __align__(128) uint4 buf[B2][NT][2];
    __shared__ __align__(8) u64 bar[NT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 barp = (u32)__cvta_generic_to_shared(&bar[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barp) : "memory");
    __syncwarp();
    const long long gw = (long long)blockIdx.x * (NT / 32) + warp, tw = (long long)gridDim.x * (NT / 32);
    u64 acc[4] = {0, 0, 0, 0};
    u32 phase = 0;
    for (long long base = gw * 32 * B2; base + 32 * B2 <= m; base += tw * 32 * B2) {
        int v[B2];
#pragma unroll
        for (int i = 0; i < B2; ++i) v[i] = idx[base + i * 32 + lane];
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barp), "r"(32u * 32u * B2) : "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < B2; ++i) {
            const u32 dst = (u32)__cvta_generic_to_shared(&buf[i][threadIdx.x][0]);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                         ::"r"(dst), "l"(tab + (size_t)v[i] * 2), "r"(barp) : "memory");
        }
        u32 ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(barp), "r"(phase) : "memory");
        phase ^= 1;
#pragma unroll
        for (int i = 0; i < B2; ++i) {
            const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(&buf[i][threadIdx.x][0]), b = *reinterpret_cast<const ulonglong2 *>(&buf[i][threadIdx.x][1]);
            acc[0] |= a.x; acc[1] |= a.y; acc[2] |= b.x; acc[3] |= b.y;
        }
        __syncwarp();
    }
    if ((acc[0] | acc[1] | acc[2] | acc[3]) == 0x12345678ull) out[0] = acc[0];
}

int main(int argc, char **argv)
{
    const int rows = argc > 1 ? atoi(argv[1]) : 89250;
    const long long m = 148ll * 768 * 256;  // gathers per launch (29 M)
    std::vector<int> h(m);
    uint64_t s = 88172645463325252ull;
    for (long long i = 0; i < m; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (uint64_t)rows); }
    uint4 *tab; int *idx; u64 *out;
    CK(cudaMalloc(&tab, (size_t)rows * 32)); CK(cudaMemset(tab, 0, (size_t)rows * 32));
    CK(cudaMalloc(&idx, m * 4)); CK(cudaMemcpy(idx, h.data(), m * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, 64));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int grid = prop.multiProcessorCount * 2;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](const char *name, auto launch) {
        for (int i = 0; i < 2; ++i) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 5; ++i) launch();
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
        const double rows_per_ns = m / (ms * 1e6);
        printf("%-34s %8.3f ms  %7.2f rows/ns  %6.2f TB/s (32 B rows)  %6.2f SM-cycles/row @1.9GHz\n", name, ms, rows_per_ns,
               rows_per_ns * 32e-3, 1.9 * prop.multiProcessorCount / rows_per_ns);
    };
    printf("table %d rows x 32 B = %.2f MB, %lld gathers, grid %d x %d\n", rows, rows * 32e-6, m, grid, NT);
    run("pair/row LDG.128 (current)", [&] { v_pair128<false><<<grid, NT>>>(tab, idx, m, out); });
    run("pair/row LDG.128 .cg", [&] { v_pair128<true><<<grid, NT>>>(tab, idx, m, out); });
    run("4 lanes/row LDG.64", [&] { v_lanes<4><<<grid, NT>>>((const u32 *)tab, idx, m, out); });
    run("8 lanes/row LDG.32", [&] { v_lanes<8><<<grid, NT>>>((const u32 *)tab, idx, m, out); });
    run("thread/row 2xLDG.128", [&] { v_thread<false><<<grid, NT>>>(tab, idx, m, out); });
    run("thread/row LDG.256", [&] { v_thread<true><<<grid, NT>>>(tab, idx, m, out); });
    run("pair/row, 1 pair active per LDG", [&] { v_single<<<grid, NT>>>(tab, idx, m, out); });
    run("pair/row LDGSTS 16 B", [&] { v_ldgsts<<<grid, NT>>>(tab, idx, m, out); });
    run("thread/row cp.async.bulk 32 B", [&] { v_bulk<<<grid, NT>>>(tab, idx, m, out); });
    return 0;
}
