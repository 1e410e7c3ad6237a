// fp32 row sweep: instantiations of sangnom_wide.cuh for float samples.
#include "sangnom_launch_wide.cuh"

namespace sn {
namespace launch {

cudaError_t launch_f32(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream) { return launch_wide<float>(tasks, ntasks, g, stream); }

}  // namespace launch
}  // namespace sn
