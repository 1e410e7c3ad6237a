// Kernel launchers of libsangnom_cuda (sm_100a). The kernels themselves live in
//   sangnom_u8.cuh    8-bit samples: packed 16-bit-lane arithmetic, 8 columns per thread
//   sangnom_wide.cuh  16-bit and fp32 samples: one pixel per lane, 4 columns per thread
// Both fuse the reference's three stages (/root/reference/src/SangNom2.cpp prepareBuffers_c :74-124,
// processBuffers_c :126-159, finalizePlane_c :161-257) into one sweep down the pool rows so that the
// nine cost buffers never exist in memory, and both split a plane that is too wide for one block
// over the blocks of a thread-block cluster (sangnom_cluster.cuh).
#include "sangnom_kernels.h"
#include "sangnom_u8.cuh"
#include "sangnom_wide.cuh"
#include "sangnom_turn.cuh"
#include "sangnom_turn_tma.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace sn {

namespace {

int env_int(const char* name, int def)
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
int cluster_split(int S, int seg_pref, int seg_hard, int cols_per_thread)
{
    int fallback = 0;
    for (int G = 1; G <= 8; G *= 2) {
        if (S % G != 0 || (S / G) % cols_per_thread != 0 || S / G > seg_hard) continue;
        if (S / G <= seg_pref) return G;
        fallback = G;
    }
    return fallback;
}

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

}  // namespace

int max_pool_width(int sample_bytes) { return sample_bytes == 1 ? 8 * 2048 : 8 * 1024; }

bool pool_width_supported(int sample_bytes, int S)
{
    if (S <= 0 || S % 32 != 0 || S > max_pool_width(sample_bytes)) return false;
    return sample_bytes == 1 ? cluster_split(S, 2048, 2048, u8k::kCols) != 0 : cluster_split(S, 1024, 1024, wide::kCols) != 0;
}

const char* kernel_variant_name(int sample_bytes, int S)
{
    (void)S;
    switch (sample_bytes) {
        case 1: return "sangnom_u8_row_sweep";
        case 2: return "sangnom_wide_row_sweep<u16>";
        default: return "sangnom_wide_row_sweep<f32>";
    }
}

namespace {

// cuTensorMapEncodeTiled, looked up at run time (no link dependency on libcuda)
using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiled encode_tiled()
{
    static const EncodeTiled fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiled>(p);
        cudaGetLastError();
        return static_cast<EncodeTiled>(nullptr);
    }();
    return fn;
}

// A plane the TMA can address: base and pitch multiples of 16 bytes (and sizes within the tensor map's limits).
bool tma_addressable(const void* p, long long pitch, int width, int height)
{
    return ((reinterpret_cast<uintptr_t>(p) | (uintptr_t)pitch) & 15) == 0 && pitch > 0 && width > 0 && height > 0 && pitch < (1ll << 40);
}

bool encode_plane_map(CUtensorMap* map, int sample_bytes, const void* base, long long pitch, int width, int height)
{
    const EncodeTiled enc = encode_tiled();
    if (!enc) return false;
    const CUtensorMapDataType type = sample_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (sample_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32);
    const cuuint64_t dims[2] = { (cuuint64_t)width, (cuuint64_t)height };
    const cuuint64_t strides[1] = { (cuuint64_t)pitch };
    const cuuint32_t side = (cuuint32_t)turn::tile_side(sample_bytes);
    const cuuint32_t box[2] = { side, side };            // 128 bytes x 32*m rows
    const cuuint32_t estr[2] = { 1, 1 };
    return enc(map, type, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int kBytes>
cudaError_t launch_turn_tma(const turn::TmaBatch& batch, int n, int tiles, int fr, int fc, cudaStream_t stream)
{
    static size_t configured[64] = {};
    auto kernel = turn::sangnom_turn_planes_tma<kBytes>;
    const size_t smem = turn::tma_smem_bytes(kBytes);
    cudaError_t e = ensure_smem(kernel, smem, configured);
    if (e != cudaSuccess) return e;
    const int blocks = (tiles + turn::kTmaTilesPerBlock - 1) / turn::kTmaTilesPerBlock;
    kernel<<<blocks, turn::kTmaThreads, smem, stream>>>(batch, n, tiles, fr, fc);
    return cudaGetLastError();
}

// the planes TMA cannot address (or all of them with SANGNOM_TURN=plain): register / shared-memory path
cudaError_t launch_turn_plain(int sample_bytes, const TurnPlane* const* planes, int nplanes, int fr, int fc, cudaStream_t stream, int* launches)
{
    const int TS = turn::tile_side(sample_bytes);
    const size_t smem = turn::smem_bytes(sample_bytes);
    for (int first = 0; first < nplanes; first += turn::kMaxTasks) {
        const int n = std::min(turn::kMaxTasks, nplanes - first);
        turn::TurnBatch batch{};
        int tiles = 0;
        for (int i = 0; i < n; ++i) {
            const TurnPlane& p = *planes[first + i];
            turn::TurnTask& t = batch.t[i];
            t.src = p.src; t.dst = p.dst; t.src_pitch = p.src_pitch; t.dst_pitch = p.dst_pitch;
            t.width = p.width; t.height = p.height;
            t.tiles_x = (p.width + TS - 1) / TS;
            t.first_block = tiles;
            tiles += t.tiles_x * ((p.height + TS - 1) / TS);
        }
        const int blocks = (tiles + turn::kTilesPerBlock - 1) / turn::kTilesPerBlock;
        switch (sample_bytes) {
            case 1: turn::sangnom_turn_planes<1><<<blocks, turn::kThreads, smem, stream>>>(batch, n, tiles, fr, fc); break;
            case 2: turn::sangnom_turn_planes<2><<<blocks, turn::kThreads, smem, stream>>>(batch, n, tiles, fr, fc); break;
            default: turn::sangnom_turn_planes<4><<<blocks, turn::kThreads, smem, stream>>>(batch, n, tiles, fr, fc); break;
        }
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (launches) ++*launches;
    }
    return cudaSuccess;
}

}  // namespace

cudaError_t launch_turn_planes(int sample_bytes, const TurnPlane* planes, int nplanes, TurnKind kind, cudaStream_t stream, int* launches)
{
    if (launches) *launches = 0;
    if (nplanes <= 0) return cudaSuccess;
    if (sample_bytes != 1 && sample_bytes != 2 && sample_bytes != 4) return cudaErrorInvalidValue;
    const int TS = turn::tile_side(sample_bytes);
    const int fr = kind == kTurnLeft, fc = kind == kTurnRight;
    for (int i = 0; i < nplanes; ++i)
        if (planes[i].width <= 0 || planes[i].height <= 0) return cudaErrorInvalidValue;
    // TMA path for every plane it can address; the rest take the plain kernel
    static const bool plain_only = [] { const char* v = getenv("SANGNOM_TURN"); return v && std::string(v) == "plain"; }();
    std::vector<const TurnPlane*> plain, viaTma;
    for (int i = 0; i < nplanes; ++i) {
        const TurnPlane& p = planes[i];
        // (a turn that mirrors the source columns tiles the plane from its right edge: the tile origins are multiples of
        // 16 bytes - which the tensor-map load needs of its inner coordinate - only if the row length is)
        const bool ok = !plain_only && encode_tiled() != nullptr && tma_addressable(p.src, p.src_pitch, p.width, p.height) &&
                        tma_addressable(p.dst, p.dst_pitch, p.height, p.width) && (!fr || ((long long)p.width * sample_bytes) % 16 == 0) &&
                        (p.dst_padding_writable || ((long long)p.height * sample_bytes) % 16 == 0);
        (ok ? viaTma : plain).push_back(&p);
    }
    for (size_t first = 0; first < viaTma.size(); first += turn::kTmaMaxPlanes) {
        const int n = (int)std::min<size_t>(turn::kTmaMaxPlanes, viaTma.size() - first);
        turn::TmaBatch batch{};
        int tiles = 0, kept = 0;
        for (int i = 0; i < n; ++i) {
            const TurnPlane& p = *viaTma[first + i];
            if (!encode_plane_map(&batch.src[kept], sample_bytes, p.src, p.src_pitch, p.width, p.height) ||
                !encode_plane_map(&batch.dst[kept], sample_bytes, p.dst, p.dst_pitch, p.height, p.width)) { plain.push_back(&p); continue; }
            turn::TmaPlane& t = batch.plane[kept++];
            t.width = p.width; t.height = p.height;
            t.tiles_x = (p.width + TS - 1) / TS;
            t.first_tile = tiles;
            tiles += t.tiles_x * ((p.height + TS - 1) / TS);
        }
        if (kept == 0) continue;
        cudaError_t e;
        switch (sample_bytes) {
            case 1: e = launch_turn_tma<1>(batch, kept, tiles, fr, fc, stream); break;
            case 2: e = launch_turn_tma<2>(batch, kept, tiles, fr, fc, stream); break;
            default: e = launch_turn_tma<4>(batch, kept, tiles, fr, fc, stream); break;
        }
        if (e != cudaSuccess) return e;
        if (launches) ++*launches;
    }
    if (!plain.empty()) return launch_turn_plain(sample_bytes, plain.data(), (int)plain.size(), fr, fc, stream, launches);
    return cudaSuccess;
}

cudaError_t launch_plane_tasks(int sample_bytes, const PlaneTask* tasks_dev, int ntasks, LaunchGeometry g, cudaStream_t stream)
{
    if (ntasks <= 0) return cudaSuccess;
    if (g.S % 32 != 0 || g.S > max_pool_width(sample_bytes)) return cudaErrorInvalidValue;
    switch (sample_bytes) {
        case 1: return launch_u8(tasks_dev, ntasks, g, stream);
        case 2: return launch_wide<uint16_t>(tasks_dev, ntasks, g, stream);
        case 4: return launch_wide<float>(tasks_dev, ntasks, g, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sn
