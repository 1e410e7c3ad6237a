// Plane transposition / quarter turns with the tensor memory accelerator (TMA), for the anti-aliasing chain
//   SangNom2(dh=true) -> turn -> SangNom2(dh=true) -> turn back          (SURVEY.md 8(f)2).
// Same function as sangnom_turn.cuh (which stays as the path for planes TMA cannot address: base or pitch not a
// multiple of 16 bytes):   out[fr ? W-1-x : x][fc ? H-1-y : y] = in[y][x].
//
// HBM-bound tile movement, so the data path is: one elected thread issues `cp.async.bulk.tensor.2d` loads of whole
// tiles (128 bytes x 32*m rows, m = samples per 32-bit word) into a ring of shared-memory stages, completion on an
// mbarrier; the block transposes the tile shared -> shared through registers (m x m sample cells, byte permutes);
// the elected thread writes the result tile back with a `cp.async.bulk.tensor.2d` store. Tiles that stick out of
// the plane need no special path: loads fill out-of-bounds samples with zeros (negative coordinates included), stores
// clip what lies beyond the extent. Both tiles use the
// 128-byte swizzle of the tensor map (the 16-byte chunk index of a row is XORed with the row index mod 8), which
// makes the row-wise reads conflict-free and the column-wise writes at most 4-way conflicted - shared memory has
// several times the bandwidth this kernel needs.
#pragma once
#include <cuda.h>          // CUtensorMap (type only)
#include <stdint.h>

#include "sangnom_turn.cuh"

namespace sn {
namespace turn {

constexpr int kTmaMaxPlanes = 32;        // planes per launch: the tensor maps travel as kernel parameters (2 x 128 B per plane)
constexpr int kTmaThreads = 256;
constexpr int kTmaStages = 3;            // tiles in flight per block
constexpr int kTmaTilesPerBlock = 16;    // consecutive tiles per block

struct TmaPlane {
    int width, height;      // W, H of the source
    int tiles_x;            // tiles per tile row
    int first_tile;         // index of the plane's first tile in the launch
};
struct alignas(64) TmaBatch {
    CUtensorMap src[kTmaMaxPlanes];      // W x H samples, box TS x TS, 128-byte swizzle
    CUtensorMap dst[kTmaMaxPlanes];      // H x W samples
    TmaPlane plane[kTmaMaxPlanes];
};

inline size_t tma_smem_bytes(int sample_bytes) { return (size_t)kTmaStages * 2 * tile_side(sample_bytes) * 128 + 1024; }

#ifndef SN_HOST_EMULATION
namespace tma {
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void load_tile(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(saddr(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void store_tile(const CUtensorMap* map, int c0, int c1, const void* smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(saddr(smem_src)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int kPending> __device__ __forceinline__ void wait_stores_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void wait_stores_done() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (the TMA store that follows)
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "TWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TDONE_%=;\n\t"
        "bra TWAIT_%=;\n"
        "TDONE_%=:\n\t}"
        ::"r"(saddr(b)), "r"(parity) : "memory");
}
}  // namespace tma

// byte offset of 32-bit word `w` (0..31) of row `r` inside a tile of 128-byte rows under the 128-byte swizzle
__device__ __forceinline__ int swz(int r, int w) { return r * 128 + ((((w >> 2) ^ (r & 7)) << 4) | ((w & 3) << 2)); }

template <int kBytes>
__global__ void __launch_bounds__(kTmaThreads)
sangnom_turn_planes_tma(const __grid_constant__ TmaBatch batch, int nplanes, int total_tiles, int flip_rows, int flip_cols)
{
    constexpr int m = 4 / kBytes;               // samples per word = cell side
    constexpr int TS = 32 * m;                  // tile side in samples = rows of a tile; a tile row is 128 bytes
    constexpr int kTileBytes = TS * 128;
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* const smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full[kTmaStages];
    auto tile_in = [&](int s) { return smem + (size_t)s * 2 * kTileBytes; };
    auto tile_out = [&](int s) { return smem + (size_t)s * 2 * kTileBytes + kTileBytes; };

    // Tile order as in sangnom_turn.cuh: bands of kBandRows tile rows, inside a band down the columns first, so that the
    // blocks running at the same time cover a patch that is several tiles wide and high.
    struct Tile { int plane, x0, y0; };
    auto locate = [&](int g) -> Tile {
        int lo = 0, hi = nplanes - 1;                                   // binary search of the plane (block-uniform)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (batch.plane[mid].first_tile <= g) lo = mid; else hi = mid - 1;
        }
        const TmaPlane& p = batch.plane[lo];
        const int tile = g - p.first_tile;
        const int tiles_y = (p.height + TS - 1) / TS;
        const int band = tile / (p.tiles_x * kBandRows), rem = tile - band * (p.tiles_x * kBandRows);
        const int rows_in_band = min(kBandRows, tiles_y - band * kBandRows);
        const int col = rem / rows_in_band, row = band * kBandRows + rem - col * rows_in_band;
        // A flipped axis is tiled from its far end, so that the ragged tile is the one whose SOURCE coordinate is
        // negative (loads fill with zeros) and every DESTINATION coordinate stays >= 0: the tensor-map store clips what
        // sticks out beyond the extent, but faults on a negative start coordinate.
        return Tile{ lo, flip_rows ? p.width - (col + 1) * TS : col * TS, flip_cols ? p.height - (row + 1) * TS : row * TS };
    };

    const int g0 = (int)blockIdx.x * kTmaTilesPerBlock, ntiles = min(kTmaTilesPerBlock, total_tiles - g0);
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) tma::mbar_init(&full[s], 1);
        tma::fence_init();
    }
    __syncthreads();
    auto issue_load = [&](int i) {
        const Tile q = locate(g0 + i);
        tma::load_tile(tile_in(i % kTmaStages), &batch.src[q.plane], q.x0, q.y0, &full[i % kTmaStages], (unsigned)kTileBytes);
    };
    if (tid == 0)
        for (int i = 0; i < kTmaStages - 1 && i < ntiles; ++i) issue_load(i);

    const int cxl = lane & 7, cyl = lane >> 3;      // a warp moves patches of 8 x 4 cells
    for (int i = 0; i < ntiles; ++i) {
        const int s = i % kTmaStages;
        if (tid == 0) tma::wait_stores_read<kTmaStages - 1>();         // the store that last read tile_out(s) has read it
        __syncthreads();                                                // ... and everybody is done with tile_in((i-1) % stages)
        if (tid == 0 && i + kTmaStages - 1 < ntiles) issue_load(i + kTmaStages - 1);
        tma::mbar_wait(&full[s], (unsigned)(i / kTmaStages) & 1u);
        const unsigned char* const A = tile_in(s);
        unsigned char* const B = tile_out(s);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int patch = warp + 8 * q;
            const int cx = (patch & 3) * 8 + cxl, cy = (patch >> 2) * 4 + cyl;      // cell: source word column cx, cell row cy
            uint32_t in[m], out[m];
#pragma unroll
            for (int k = 0; k < m; ++k) in[k] = *reinterpret_cast<const uint32_t*>(A + swz(m * cy + k, cx));
            cell_transpose<kBytes>(in, out);
            const int bcol = flip_cols ? 31 - cy : cy;
#pragma unroll
            for (int j = 0; j < m; ++j) {
                const int brow = flip_rows ? TS - 1 - (m * cx + j) : m * cx + j;
                *reinterpret_cast<uint32_t*>(B + swz(brow, bcol)) = flip_cols ? reverse_samples<kBytes>(out[j]) : out[j];
            }
        }
        tma::fence_async_shared();
        __syncthreads();
        if (tid == 0) {
            const Tile q = locate(g0 + i);
            const TmaPlane& p = batch.plane[q.plane];
            const int dcol = flip_cols ? p.height - (q.y0 + TS) : q.y0;      // >= 0 by the choice of the tile origins
            const int drow = flip_rows ? p.width - (q.x0 + TS) : q.x0;
            tma::store_tile(&batch.dst[q.plane], dcol, drow, B);
        }
    }
    if (tid == 0) tma::wait_stores_done();
}
#endif  // SN_HOST_EMULATION

}  // namespace turn
}  // namespace sn
