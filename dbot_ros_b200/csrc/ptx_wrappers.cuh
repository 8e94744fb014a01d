/*
 * ptx_wrappers.cuh -- inline-PTX wrappers: mbarrier, TMA bulk copies, named barriers, 256-bit streaming loads/stores
 * Part of libdbot_b200.so; included by dbot_b200.cu (one translation unit).
 */
#pragma once

/* ------------------------------------------------------------------------------------------ */
/* PTX wrappers: mbarrier, TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP), named barriers,     */
/* 256-bit streaming loads / stores (SASS LDG.256 / STG.256, new with sm_100)                   */
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
/* barrier among the compute warps only (the streaming warps never take part) */
__device__ __forceinline__ void compute_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(DBOT_CT) : "memory");
}
/* barrier among the streaming warps only */
__device__ __forceinline__ void stream_sync()
{
    asm volatile("bar.sync 2, %0;" ::"n"(DBOT_ST) : "memory");
}
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8])
{
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]),
                   "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256_stream(float* p, const float (&v)[8])
{
    asm volatile(
        "st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
        : "memory");
}

