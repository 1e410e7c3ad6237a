// Plane transposition / quarter turns on the device, for the anti-aliasing chain
//   SangNom2(dh=true) -> turn -> SangNom2(dh=true) -> turn back          
// so that a frame stays in HBM between the two interpolation passes instead of crossing PCIe three more times.
// There is no counterpart in /root/reference: there the turns are separate host filters of the script
// (TurnLeft/TurnRight in AviSynth, std.Transpose in VapourSynth, which SangNom2 was ported from - README.md:5).
//
// out[fr ? W-1-x : x][fc ? H-1-y : y] = in[y][x]       (out has H columns and W rows)
//   (fr, fc) = (0,0) transpose, (0,1) TurnRight (clockwise), (1,0) TurnLeft
//
// HBM-bound: every sample is read once and written once. A tile is 32x32 CELLS, a cell being m x m samples with
// m = 4 / sizeof(sample), i.e. one 32-bit word per cell row; a thread loads the m words of a cell (a warp reads 128
// contiguous bytes of a row), transposes the cell in registers with byte permutes, and the 32x32 word transposition
// goes through shared memory ([32][33] words per word plane, conflict-free both ways). Tiles that stick out of the
// plane, or planes whose rows are not word-aligned, take a sample-by-sample path.
#pragma once
#include <stdint.h>

#ifndef SN_DYNAMIC_SMEM
#define SN_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace sn {
namespace turn {

struct TurnTask {
    const void* src;        // W x H samples
    void* dst;              // H x W samples
    long long src_pitch;    // bytes
    long long dst_pitch;    // bytes
    int width, height;      // W, H of the source
    int tiles_x;            // tiles per tile row of this task
    int first_block;        // index of the task's first tile in the launch
};

constexpr int kThreads = 256;           // 32 x 8
inline __host__ __device__ int tile_side(int sample_bytes) { return 32 * (4 / sample_bytes); }     // samples
inline size_t smem_bytes(int sample_bytes) { return (size_t)(4 / sample_bytes) * 32 * 33 * 4; }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }

// m x m cell transposition: in[k] = word of cell row k (samples left to right in increasing byte order);
// out[j] = word holding sample j of rows 0..m-1.
template <int kBytes> __device__ __forceinline__ void cell_transpose(const uint32_t (&in)[4 / kBytes], uint32_t (&out)[4 / kBytes])
{
    if constexpr (kBytes == 4) {
        out[0] = in[0];
    } else if constexpr (kBytes == 2) {
        out[0] = prmt(in[0], in[1], 0x5410);
        out[1] = prmt(in[0], in[1], 0x7632);
    } else {
        const uint32_t a = prmt(in[0], in[1], 0x5140), b = prmt(in[0], in[1], 0x7362);     // rows 0,1 interleaved: bytes (0,0)(1,0)(0,1)(1,1) / (0,2)(1,2)(0,3)(1,3)
        const uint32_t c = prmt(in[2], in[3], 0x5140), d = prmt(in[2], in[3], 0x7362);
        out[0] = prmt(a, c, 0x5410);
        out[1] = prmt(a, c, 0x7632);
        out[2] = prmt(b, d, 0x5410);
        out[3] = prmt(b, d, 0x7632);
    }
}

// reverse the order of the samples inside a word
template <int kBytes> __device__ __forceinline__ uint32_t reverse_samples(uint32_t w)
{
    if constexpr (kBytes == 4) return w;
    else if constexpr (kBytes == 2) return prmt(w, 0u, 0x1032);
    else return prmt(w, 0u, 0x0123);
}

// One launch serves many planes: block b belongs to the task whose [first_block, first_block + tiles) range holds it.
template <int kBytes>
__global__ void __launch_bounds__(kThreads)
sangnom_turn_planes(const TurnTask* __restrict__ tasks, int ntasks, int flip_rows, int flip_cols)
{
    constexpr int m = 4 / kBytes;               // samples per word = cell side
    constexpr int TS = 32 * m;                  // tile side in samples
    SN_DYNAMIC_SMEM(smem_raw);
    uint32_t* const sm = reinterpret_cast<uint32_t*>(smem_raw);       // [m][32][33]

    // binary search of the task (block-uniform)
    int lo = 0, hi = ntasks - 1;
    const int b = (int)blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tasks[mid].first_block <= b) lo = mid; else hi = mid - 1;
    }
    const TurnTask t = tasks[lo];
    const int tile = b - t.first_block;
    const int x0 = (tile % t.tiles_x) * TS, y0 = (tile / t.tiles_x) * TS;      // tile origin in the source
    const int W = t.width, H = t.height;
    const int tx = (int)threadIdx.x & 31, ty = (int)threadIdx.x >> 5;
    const unsigned char* const src = static_cast<const unsigned char*>(t.src);
    unsigned char* const dst = static_cast<unsigned char*>(t.dst);

    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)t.src_pitch | (uintptr_t)t.dst_pitch) & 3) == 0 &&
                         (!flip_cols || H % m == 0);
    if (aligned && x0 + TS <= W && y0 + TS <= H) {
        // ---- fast path: whole tile, word accesses ----
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ci = ty + 8 * i;                                  // cell row inside the tile
            uint32_t in[m], out[m];
#pragma unroll
            for (int k = 0; k < m; ++k)
                in[k] = *reinterpret_cast<const uint32_t*>(src + (long long)(y0 + ci * m + k) * t.src_pitch + (long long)(x0 + tx * m) * kBytes);
            cell_transpose<kBytes>(in, out);
#pragma unroll
            for (int j = 0; j < m; ++j) sm[(j * 32 + tx) * 33 + ci] = out[j];      // [j][cell col][cell row]
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int oi = ty + 8 * i;                                  // source cell column = output cell row
#pragma unroll
            for (int j = 0; j < m; ++j) {
                uint32_t w = sm[(j * 32 + oi) * 33 + tx];               // samples (x, y .. y+m-1), x = x0 + oi*m + j, y = y0 + tx*m
                const int x = x0 + oi * m + j, y = y0 + tx * m;
                const int orow = flip_rows ? W - 1 - x : x;
                int ocol = y;
                if (flip_cols) { ocol = H - y - m; w = reverse_samples<kBytes>(w); }
                *reinterpret_cast<uint32_t*>(dst + (long long)orow * t.dst_pitch + (long long)ocol * kBytes) = w;
            }
        }
        return;
    }
    // ---- edge tiles / unaligned planes: sample by sample ----
    for (int e = (int)threadIdx.x; e < TS * TS; e += kThreads) {
        const int ly = e / TS, lxx = e % TS;                            // consecutive threads: consecutive source columns
        const int x = x0 + lxx, y = y0 + ly;
        if (x >= W || y >= H) continue;
        const int orow = flip_rows ? W - 1 - x : x, ocol = flip_cols ? H - 1 - y : y;
        const unsigned char* s = src + (long long)y * t.src_pitch + (long long)x * kBytes;
        unsigned char* d = dst + (long long)orow * t.dst_pitch + (long long)ocol * kBytes;
        if constexpr (kBytes == 1) *d = *s;
        else if constexpr (kBytes == 2) *reinterpret_cast<uint16_t*>(d) = *reinterpret_cast<const uint16_t*>(s);
        else *reinterpret_cast<uint32_t*>(d) = *reinterpret_cast<const uint32_t*>(s);
    }
}

}  // namespace turn
}  // namespace sn
