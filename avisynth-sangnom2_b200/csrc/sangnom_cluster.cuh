// Thread-block-cluster helpers: a plane wider than one block can hold is split into column
// segments, one block each, the blocks of a plane forming one cluster. Per pool row the edge
// threads push their 3-column halo of the vertical sums straight into the neighbour block's
// shared memory (DSMEM) with asynchronous stores that complete on an mbarrier of the RECEIVING
// block (st.async ... mbarrier::complete_tx, SASS STAS): the data's arrival is the signal. Only the
// one edge thread that reads a halo waits for it; there is no cluster-wide barrier and no memory
// fence per row (barrier.cluster.arrive.release costs a MEMBAR.ALL.GPU, i.e. every warp waits for
// its outstanding picture-row stores to be acknowledged, once per row). Neighbouring blocks stay
// within one row of each other by data dependence alone: the thread that consumes the halo of
// row r is the thread that sends the opposite halo of row r+1, and a halo of row r+2 is sent
// only after that one has been received.
#pragma once
#include <cstdint>

namespace sn {
namespace cl {

#ifdef SN_HOST_EMULATION
inline unsigned rank() { return emul::crank; }
inline unsigned size() { return emul::csize; }
inline void sync_all() { emul::cluster_bar->arrive_and_wait(); }
template <typename V>
inline void store_remote(V* local_ptr, unsigned target_rank, V value)
{
    const size_t off = reinterpret_cast<unsigned char*>(local_ptr) - emul::smem;
    *reinterpret_cast<V*>(emul::cluster_smem[target_rank] + off) = value;
}
inline void store_remote4(uint4* local_ptr, unsigned target_rank, uint4 value) { store_remote(local_ptr, target_rank, value); }
// Halo barrier: `rx` counts the bytes that have arrived, `taken` the bytes the (single) waiting thread has consumed.
struct HaloBar { uint64_t rx, taken; };
inline void halo_init(HaloBar* b) { __atomic_store_n(&b->rx, (uint64_t)0, __ATOMIC_RELEASE); b->taken = 0; }
inline void halo_fence_init() {}
template <typename V>
inline void store_remote_tx(V* local_ptr, unsigned target_rank, V value, HaloBar* local_bar)
{
    store_remote(local_ptr, target_rank, value);
    const size_t off = reinterpret_cast<unsigned char*>(local_bar) - emul::smem;
    __atomic_fetch_add(&reinterpret_cast<HaloBar*>(emul::cluster_smem[target_rank] + off)->rx, (uint64_t)sizeof(V), __ATOMIC_RELEASE);
}
inline void halo_wait(HaloBar* b, unsigned, unsigned bytes)
{
    while (__atomic_load_n(&b->rx, __ATOMIC_ACQUIRE) < b->taken + bytes) std::this_thread::yield();
    b->taken += bytes;
}
#else
__device__ __forceinline__ unsigned rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// all threads of all blocks of the cluster; release/acquire orders the DSMEM stores before it
__device__ __forceinline__ void sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store `value` at the address that `local_ptr` has in block `target_rank`'s shared memory
template <typename V>
__device__ __forceinline__ void store_remote(V* local_ptr, unsigned target_rank, V value)
{
    static_assert(sizeof(V) == 4 || sizeof(V) == 2, "32- or 16-bit stores only");
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(local_ptr);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(target_rank));
    if constexpr (sizeof(V) == 4) {
        uint32_t bits;
        memcpy(&bits, &value, 4);
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(bits) : "memory");
    } else {
        uint16_t bits;
        memcpy(&bits, &value, 2);
        asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(remote), "h"(bits) : "memory");
    }
}
// the same for one 16-byte vector (16-byte aligned)
__device__ __forceinline__ void store_remote4(uint4* local_ptr, unsigned target_rank, uint4 value)
{
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(local_ptr);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(target_rank));
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "r"(value.x), "r"(value.y), "r"(value.z), "r"(value.w) : "memory");
}
// Halo barrier: an mbarrier in the RECEIVING block's shared memory, one arrival (the waiting thread's own, which also
// announces the bytes) plus the bytes of the neighbour's asynchronous stores per phase.
struct __align__(8) HaloBar { unsigned long long v; };
__device__ __forceinline__ void halo_init(HaloBar* b)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
// make the initialised barriers visible to the other blocks of the cluster (before the one cluster barrier at start)
__device__ __forceinline__ void halo_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// store `value` at the address `local_ptr` has in block `target_rank`; completes on that block's copy of `local_bar`
template <typename V>
__device__ __forceinline__ void store_remote_tx(V* local_ptr, unsigned target_rank, V value, HaloBar* local_bar)
{
    static_assert(sizeof(V) == 16 || sizeof(V) == 8, "uint4 or uint2");
    uint32_t dst, bar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"((uint32_t)__cvta_generic_to_shared(local_ptr)), "r"(target_rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar) : "r"((uint32_t)__cvta_generic_to_shared(local_bar)), "r"(target_rank));
    if constexpr (sizeof(V) == 16)
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                     ::"r"(dst), "r"(value.x), "r"(value.y), "r"(value.z), "r"(value.w), "r"(bar) : "memory");
    else
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                     ::"r"(dst), "r"(value.x), "r"(value.y), "r"(bar) : "memory");
}
// one thread: announce `bytes` for the current phase, arrive, and wait until the phase with this parity is complete
__device__ __forceinline__ void halo_wait(HaloBar* b, unsigned parity, unsigned bytes)
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "HWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HDONE_%=;\n\t"
        "bra HWAIT_%=;\n"
        "HDONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
#endif

}  // namespace cl
}  // namespace sn
