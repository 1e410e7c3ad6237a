// 16-bit row sweep: instantiations of sangnom_wide.cuh for uint16_t samples.
#include "sangnom_launch_wide.cuh"

namespace sn {
namespace launch {

cudaError_t launch_u16(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream) { return launch_wide<uint16_t>(tasks, ntasks, g, stream); }
bool wide_width_supported(int S) { return cluster_split(S, 1024, 1024, wide::kCols) != 0; }

}  // namespace launch
}  // namespace sn
