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

#ifndef SN_GRID_CONSTANT
#ifdef SN_HOST_EMULATION
#define SN_GRID_CONSTANT
#else
#define SN_GRID_CONSTANT __grid_constant__
#endif
#endif
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

constexpr int kMaxTasks = 64;           // planes per launch: the task table travels as a kernel parameter (3 KB)
struct TurnBatch { TurnTask t[kMaxTasks]; };

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

constexpr int kBandRows = 16;
constexpr int kTilesPerBlock = 8;        // consecutive tiles per block: the next tile's loads are in flight while this one is written

// One launch serves many planes: tile index g belongs to the task whose [first_block, first_block + tiles) range holds it.
template <int kBytes>
__global__ void __launch_bounds__(kThreads)
sangnom_turn_planes(const SN_GRID_CONSTANT TurnBatch batch, int ntasks, int total_tiles, int flip_rows, int flip_cols)
{
    const TurnTask* const tasks = batch.t;
    constexpr int m = 4 / kBytes;               // samples per word = cell side
    constexpr int TS = 32 * m;                  // tile side in samples
    SN_DYNAMIC_SMEM(smem_raw);
    uint32_t* const sm = reinterpret_cast<uint32_t*>(smem_raw);       // [m][32][33]
    const int tx = (int)threadIdx.x & 31, ty = (int)threadIdx.x >> 5;

    struct Tile { const unsigned char* src; unsigned char* dst; long long sp, dp; int x0, y0, W, H; bool fast;
                  int task, tiles_x, band_row0, rows_in_band, col, row; bool aligned; };
    auto finish = [&](Tile& q) {
        q.x0 = q.col * TS; q.y0 = (q.band_row0 + q.row) * TS;          // tile origin in the source
        q.fast = q.aligned && q.x0 + TS <= q.W && q.y0 + TS <= q.H;
    };
    // Tile order: bands of kBandRows tile rows, inside a band down the columns first. The few hundred blocks that
    // run at the same time then cover a patch that is many tiles wide AND high, so both the reads (contiguous
    // along source rows) and the writes (contiguous along source columns) reach DRAM as runs of a few KB instead
    // of isolated 128-byte pieces.
    auto locate = [&](int g) -> Tile {
        int lo = 0, hi = ntasks - 1;                                    // binary search of the task (block-uniform)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tasks[mid].first_block <= g) lo = mid; else hi = mid - 1;
        }
        const TurnTask& t = tasks[lo];
        Tile q;
        const int tile = g - t.first_block;
        q.task = lo; q.tiles_x = t.tiles_x;
        q.src = static_cast<const unsigned char*>(t.src); q.dst = static_cast<unsigned char*>(t.dst);
        q.sp = t.src_pitch; q.dp = t.dst_pitch; q.W = t.width; q.H = t.height;
        const int tiles_y = (t.height + TS - 1) / TS;
        const int band = tile / (t.tiles_x * kBandRows), rem = tile - band * (t.tiles_x * kBandRows);
        q.band_row0 = band * kBandRows;
        q.rows_in_band = min(kBandRows, tiles_y - q.band_row0);
        q.col = rem / q.rows_in_band; q.row = rem - q.col * q.rows_in_band;
        q.aligned = ((reinterpret_cast<uintptr_t>(q.src) | reinterpret_cast<uintptr_t>(q.dst) | (uintptr_t)q.sp | (uintptr_t)q.dp) & 3) == 0 &&
                    (!flip_cols || q.H % m == 0);
        finish(q);
        return q;
    };
    // the tile after q in the same order: one step down the column, or the top of the next column; crossing into
    // another band or plane takes the full lookup
    auto next_tile = [&](const Tile& q, int g) -> Tile {
        Tile r = q;
        if (++r.row == r.rows_in_band) { r.row = 0; if (++r.col == r.tiles_x) return locate(g); }
        finish(r);
        return r;
    };
    // the 4 cells (m words each) this thread moves of a whole tile
    auto load_cells = [&](const Tile& q, uint32_t (&r)[4][m]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < m; ++k)
                r[i][k] = __ldg(reinterpret_cast<const uint32_t*>(q.src + (long long)(q.y0 + (ty + 8 * i) * m + k) * q.sp + (long long)(q.x0 + tx * m) * kBytes));
    };

    const int g0 = (int)blockIdx.x * kTilesPerBlock, g1 = min(g0 + kTilesPerBlock, total_tiles);
    Tile cur = locate(g0);
    uint32_t regs[4][m];
    if (cur.fast) load_cells(cur, regs);
    for (int g = g0; g < g1; ++g) {
        Tile nxt = cur;
        uint32_t ahead[4][m];
        const bool more = g + 1 < g1;
        if (more) { nxt = next_tile(cur, g + 1); if (nxt.fast) load_cells(nxt, ahead); }
        if (cur.fast) {
            // ---- whole tile, word accesses ----
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t out[m];
                cell_transpose<kBytes>(regs[i], out);
#pragma unroll
                for (int j = 0; j < m; ++j) sm[(j * 32 + tx) * 33 + ty + 8 * i] = out[j];      // [j][cell col][cell row]
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int oi = ty + 8 * i;                              // source cell column = output cell row
#pragma unroll
                for (int j = 0; j < m; ++j) {
                    uint32_t w = sm[(j * 32 + oi) * 33 + tx];           // samples (x, y .. y+m-1), x = x0 + oi*m + j, y = y0 + tx*m
                    const int x = cur.x0 + oi * m + j, y = cur.y0 + tx * m;
                    const int orow = flip_rows ? cur.W - 1 - x : x;
                    int ocol = y;
                    if (flip_cols) { ocol = cur.H - y - m; w = reverse_samples<kBytes>(w); }
                    *reinterpret_cast<uint32_t*>(cur.dst + (long long)orow * cur.dp + (long long)ocol * kBytes) = w;
                }
            }
            __syncthreads();
        } else {
            // ---- edge tiles / unaligned planes: sample by sample ----
            for (int e = (int)threadIdx.x; e < TS * TS; e += kThreads) {
                const int x = cur.x0 + e % TS, y = cur.y0 + e / TS;     // consecutive threads: consecutive source columns
                if (x >= cur.W || y >= cur.H) continue;
                const int orow = flip_rows ? cur.W - 1 - x : x, ocol = flip_cols ? cur.H - 1 - y : y;
                const unsigned char* sp = cur.src + (long long)y * cur.sp + (long long)x * kBytes;
                unsigned char* dp = cur.dst + (long long)orow * cur.dp + (long long)ocol * kBytes;
                if constexpr (kBytes == 1) *dp = *sp;
                else if constexpr (kBytes == 2) *reinterpret_cast<uint16_t*>(dp) = *reinterpret_cast<const uint16_t*>(sp);
                else *reinterpret_cast<uint32_t*>(dp) = *reinterpret_cast<const uint32_t*>(sp);
            }
        }
        if (more) {
            cur = nxt;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < m; ++k) regs[i][k] = ahead[i][k];
        }
    }
}

}  // namespace turn
}  // namespace sn
