// wx_tma.cuh -- thin inline-PTX wrappers for TMA (cp.async.bulk.tensor), mbarrier and the async-proxy fences on
// sm_100a, plus host-side creation of "row maps": a buffer viewed as a 2-D tensor of 128-byte rows with the
// SWIZZLE_128B shared-memory layout (identical to wx_swz_chunk).
#pragma once
#include <cuda.h>
#include "wx_common.cuh"

__device__ __forceinline__ unsigned wx_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void wx_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wx_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void wx_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void wx_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wx_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wx_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wx_mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(wx_smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void wx_tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     wx_smem_u32(smem_dst)),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(wx_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void wx_tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<unsigned long long>(map)),
                 "r"(c0), "r"(c1), "r"(wx_smem_u32(smem_src))
                 : "memory");
}
// the same with an L2 eviction policy (createpolicy): the packet table is written once and never re-read by this kernel, x is read once
__device__ __forceinline__ unsigned long long wx_policy_evict_first()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void wx_tma_load_2d_hint(void *smem_dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar, unsigned long long pol)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
                     wx_smem_u32(smem_dst)),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(wx_smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void wx_tma_store_2d_hint(const CUtensorMap *map, int c0, int c1, const void *smem_src, unsigned long long pol)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(
                     reinterpret_cast<unsigned long long>(map)),
                 "r"(c0), "r"(c1), "r"(wx_smem_u32(smem_src)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void wx_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wx_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void wx_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void wx_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// host: tensor map over `rows` rows of 128 bytes starting at `base`, box = boxrows x 128 B, SWIZZLE_128B.
// The driver entry point is looked up at run time so the library does not link against libcuda.
int wx_make_rowmap(CUtensorMap *map, const void *base, size_t elt, long rows, long boxrows);

// 1-D bulk copies (no tensor map): 16-byte aligned addresses, size a multiple of 16 bytes.
__device__ __forceinline__ void wx_bulk_load_1d(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(wx_smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(wx_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void wx_bulk_store_1d(void *gdst, const void *smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(wx_smem_u32(smem_src)), "r"(bytes) : "memory");
}
