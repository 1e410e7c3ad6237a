// Staging of kept picture rows into shared memory: bulk asynchronous copies (TMA, `cp.async.bulk`,
// SASS UBLKCP) completing on an mbarrier. One elected thread of a block issues one copy per picture
// row a few rows ahead of the sweep; every thread waits on the row's mbarrier phase before reading.
// The source may be device memory or mapped pinned host memory (the host path reads frames straight
// over PCIe this way), so nothing here assumes HBM latency.
#pragma once
#include <cstdint>

namespace sn {
namespace stage {

#ifdef SN_HOST_EMULATION
// Emulation: the copy is done synchronously by the issuing thread, which then completes the
// barrier phase; `v` counts completed phases, so "phase with parity p done" <=> (v & 1) != p.
struct Mbar { uint64_t v; };
inline void mbar_init(Mbar* b, unsigned) { __atomic_store_n(&b->v, (uint64_t)0, __ATOMIC_RELEASE); }
inline void fence_mbar_init() {}
inline void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, Mbar* b)
{
    memcpy(smem_dst, gsrc, bytes);
    __atomic_fetch_add(&b->v, (uint64_t)1, __ATOMIC_RELEASE);
}
inline void mbar_wait(Mbar* b, unsigned parity)
{
    while ((__atomic_load_n(&b->v, __ATOMIC_ACQUIRE) & 1u) == parity) std::this_thread::yield();
}
inline void fence_generic_to_async() {}

#else
struct __align__(8) Mbar { unsigned long long v; };

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(Mbar* b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy before the first bulk copy
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one thread: announce `bytes` on the barrier, then start the copy global -> shared (16-byte aligned, bytes % 16 == 0)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, Mbar* b)
{
    const uint32_t bar = smem_addr(b);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
// order this thread's ordinary shared-memory stores before later bulk copies into the same bytes (issued by another
// thread after a block barrier)
__device__ __forceinline__ void fence_generic_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(Mbar* b, unsigned parity)
{
    const uint32_t bar = smem_addr(b);
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
#endif

}  // namespace stage
}  // namespace sn
