// Thread-block-cluster helpers: a plane wider than one block can hold is split into column
// segments, one block each, the blocks of a plane forming one cluster. Per pool row the edge
// threads push their 3-column halo of the vertical sums straight into the neighbour block's
// shared memory (DSMEM) and the whole cluster meets at one cluster barrier.
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
#endif

}  // namespace cl
}  // namespace sn
