// gp_tma.cuh — mbarrier and bulk-copy-engine (TMA, cp.async.bulk) helpers shared by gp_xcopy.cu and gp_exchange.cu.
#pragma once

#include "gp_common.cuh"

#ifdef __CUDACC__

__device__ __forceinline__ u32 gp_smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void gp_mbar_init(u32 bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void gp_mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void gp_mbar_expect_tx(u32 bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void gp_mbar_arrive(u32 bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void gp_mbar_wait(u32 bar, u32 parity)
{
    u32 done;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
    }
}

// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void gp_bulk_load(u32 dst_smem, const void *src, u32 bytes, u32 bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void gp_bulk_load_hint(u32 dst_smem, const void *src, u32 bytes, u32 bar, u64 policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}

// shared -> global, completion by bulk group
__device__ __forceinline__ void gp_bulk_store_hint(void *dst, u32 src_smem, u32 bytes, u64 policy)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
                 "r"(src_smem), "r"(bytes), "l"(policy)
                 : "memory");
}

#endif  // __CUDACC__
