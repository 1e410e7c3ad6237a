// Kernel launchers of libsangnom_cuda (sm_100a). The row-sweep kernels live in
//   sangnom_u8.cuh    8-bit samples: packed 16-bit-lane arithmetic, 8 columns per thread   (compiled in sangnom_kernels_u8.cu)
//   sangnom_wide.cuh  16-bit and fp32 samples: one pixel per lane, 4 columns per thread    (sangnom_kernels_u16.cu, _f32.cu)
// Both fuse the reference's three stages (/root/reference/src/SangNom2.cpp prepareBuffers_c :74-124,
// processBuffers_c :126-159, finalizePlane_c :161-257) into one sweep down the pool rows so that the
// nine cost buffers never exist in memory, and both split a plane that is too wide for one block
// over the blocks of a thread-block cluster (sangnom_cluster.cuh).
#include "sangnom_kernels.h"
#include "sangnom_launch.h"
#include "sangnom_turn.cuh"
#include "sangnom_turn_tma.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace sn {

using namespace launch;

int max_pool_width(int sample_bytes) { return sample_bytes == 1 ? 8 * 2048 : 8 * 1024; }

bool pool_width_supported(int sample_bytes, int S)
{
    if (S <= 0 || S % 32 != 0 || S > max_pool_width(sample_bytes)) return false;
    return sample_bytes == 1 ? u8_width_supported(S) : wide_width_supported(S);
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
        case 2: return launch_u16(tasks_dev, ntasks, g, stream);
        case 4: return launch_f32(tasks_dev, ntasks, g, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sn
