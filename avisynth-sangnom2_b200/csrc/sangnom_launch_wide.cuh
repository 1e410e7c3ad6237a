// 16-bit / fp32 row sweep (sangnom_wide.cuh): launcher template shared by sangnom_kernels_u16.cu and _f32.cu.
#pragma once
#include "sangnom_launch.h"
#include "sangnom_wide.cuh"

namespace sn {
namespace launch {

template <typename T, bool kClustered, bool kSat, bool kSpare>
cudaError_t launch_wide_variant(const PlaneTask* tasks, int ntasks, LaunchGeometry g, int G, int seg, cudaStream_t stream)
{
    static size_t configured[64] = {};
    auto kernel = wide::sangnom_wide_row_sweep<T, 256, 2, kClustered, kSat, kSpare>;
    const size_t smem = wide::smem_bytes<T>(seg);
    cudaError_t e = ensure_smem(kernel, smem, configured);
    if (e != cudaSuccess) return e;
    // kSpare: as for the 8-bit kernel - spare threads so that the last pixel thread of a narrow plane ends a warp
    const int T_ = seg / wide::kCols;
    const int threads = kSpare ? std::max(T_, std::min(256, ((T_ + 31) & ~31) + 32)) : T_;
    return launch_clustered(kernel, ntasks * G, threads, smem, G, stream, tasks, g, seg);
}

// 16-bit / fp32: 4 columns per thread, at most 1024 columns per block. fp32 has one flavour (the SSE2 path computes the
// same floats).
template <typename T>
cudaError_t launch_wide(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream)
{
    const int seg_max = std::min(std::max(env_int("SANGNOM_WIDE_SEG", 512), 128), 1024);   // tuning / test knob, read per launch
    const int G = cluster_split(g.S, seg_max, 1024, wide::kCols);
    if (G == 0) return cudaErrorInvalidValue;
    const int seg = g.S / G;
    constexpr bool kInt = !Flavour<T>::kFloat;
    g.key_mask = (unsigned)Flavour<T>::kMask << 4;                  // the key mask of the integer flavour (sangnom_wide.cuh)
    const bool spare = G == 1 && g.narrow && seg / wide::kCols < 256;
    if constexpr (kInt) {
        if (g.saturate) {
            if (G != 1) return launch_wide_variant<T, true, true, false>(tasks, ntasks, g, G, seg, stream);
            return spare ? launch_wide_variant<T, false, true, true>(tasks, ntasks, g, G, seg, stream) : launch_wide_variant<T, false, true, false>(tasks, ntasks, g, G, seg, stream);
        }
    }
    if (G != 1) return launch_wide_variant<T, true, false, false>(tasks, ntasks, g, G, seg, stream);
    return spare ? launch_wide_variant<T, false, false, true>(tasks, ntasks, g, G, seg, stream) : launch_wide_variant<T, false, false, false>(tasks, ntasks, g, G, seg, stream);
}


}  // namespace launch
}  // namespace sn
