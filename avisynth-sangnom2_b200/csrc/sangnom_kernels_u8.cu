// 8-bit row sweep (sangnom_u8.cuh): its instantiations and their launcher.
#include "sangnom_launch.h"
#include "sangnom_u8.cuh"

namespace sn {
namespace launch {

namespace {

// One instantiation per (split over a cluster?, arithmetic flavour, spare threads?); shared-memory opt-in remembered per device.
template <bool kClustered, bool kSat, bool kSpare>
cudaError_t launch_u8_variant(const PlaneTask* tasks, int ntasks, LaunchGeometry g, int G, int seg, cudaStream_t stream)
{
    static size_t configured[64] = {};
    auto kernel = u8k::sangnom_u8_row_sweep<256, 2, kClustered, kSat, kSpare>;
    const size_t smem = u8k::smem_bytes(seg);
    cudaError_t e = ensure_smem(kernel, smem, configured);
    if (e != cudaSuccess) return e;
    // kSpare: spare threads up to a whole number of warps plus one warp (at most 256); the kernel leaves the first few
    // idle so that the last pixel thread ends a warp and the state-only threads start the next one (sangnom_u8.cuh)
    const int T = seg / u8k::kCols;
    const int threads = kSpare ? std::max(T, std::min(256, ((T + 31) & ~31) + 32)) : T;
    return launch_clustered(kernel, ntasks * G, threads, smem, G, stream, tasks, g, seg);
}

// 8-bit: 8 columns per thread, at most 2048 columns per block.
}  // namespace

cudaError_t launch_u8(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream)
{
    const int seg_max = std::min(std::max(env_int("SANGNOM_U8_SEG", 2048), 256), 2048);     // tuning / test knob, read per launch: small values force cluster splits at small sizes
    const int G = cluster_split(g.S, seg_max, 2048, u8k::kCols);
    if (G == 0) return cudaErrorInvalidValue;
    const int seg = g.S / G;
    if (G == 1) {
        if (g.narrow && seg / u8k::kCols < 256)
            return g.saturate ? launch_u8_variant<false, true, true>(tasks, ntasks, g, G, seg, stream) : launch_u8_variant<false, false, true>(tasks, ntasks, g, G, seg, stream);
        return g.saturate ? launch_u8_variant<false, true, false>(tasks, ntasks, g, G, seg, stream) : launch_u8_variant<false, false, false>(tasks, ntasks, g, G, seg, stream);
    }
    return g.saturate ? launch_u8_variant<true, true, false>(tasks, ntasks, g, G, seg, stream) : launch_u8_variant<true, false, false>(tasks, ntasks, g, G, seg, stream);
}

bool u8_width_supported(int S) { return cluster_split(S, 2048, 2048, u8k::kCols) != 0; }

}  // namespace launch
}  // namespace sn
