// Internal to libsangnom_cuda: what the kernel translation units share (one per sample flavour, so that a change to
// one kernel recompiles only that kernel): launch helpers and the per-flavour launchers sangnom_kernels.cu dispatches to.
#pragma once
#include "sangnom_kernels.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

namespace sn {
namespace launch {

inline int env_int(const char* name, int def)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : def;
}

// Opt a kernel into `smem` bytes of dynamic shared memory once per device.
template <typename K>
cudaError_t ensure_smem(K kernel, size_t smem, size_t (&configured)[64])
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || smem > configured[dev]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = smem;
    }
    return cudaSuccess;
}

template <typename K, typename... Args>
cudaError_t launch_clustered(K kernel, int blocks, int threads, size_t smem, int cluster, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)cluster;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = cluster > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// Split of a pool row over the blocks of a cluster: the smallest power of two (<= 8 blocks) that brings a segment down
// to seg_pref columns with whole threads (cols_per_thread columns each); if that does not exist, the largest split
// whose segments are whole threads and fit a block (seg_hard). 0 = this pool width cannot run.
inline int cluster_split(int S, int seg_pref, int seg_hard, int cols_per_thread)
{
    int fallback = 0;
    for (int G = 1; G <= 8; G *= 2) {
        if (S % G != 0 || (S / G) % cols_per_thread != 0 || S / G > seg_hard) continue;
        if (S / G <= seg_pref) return G;
        fallback = G;
    }
    return fallback;
}

// sangnom_kernels_u8.cu / _u16.cu / _f32.cu
cudaError_t launch_u8(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream);
cudaError_t launch_u16(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream);
cudaError_t launch_f32(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream);
bool u8_width_supported(int S);
bool wide_width_supported(int S);

}  // namespace launch
}  // namespace sn
