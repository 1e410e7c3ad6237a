// Internal interface between the C-ABI layer (sangnom_api.cu) and the kernels (sangnom_kernels.cu and the per-flavour sangnom_kernels_*.cu).
#pragma once
#ifndef SN_HOST_EMULATION
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace sn {

constexpr int kNumCost = 9;   // the nine direction-cost buffers (reference SangNom2.h:8-24)

// Cost state handed from one plane pass of a frame to the next. It stands in for what the
// reference's shared, never-cleared scratch pool holds outside the next plane's rectangle
// (reference SangNom2.cpp:79-81 writes only rows 1..h/2-1, cols < w; :133-136 re-blurs all of it).
//   region A: pool rows 1..a_rows, pool columns [a_x0, S)      stored [9][a_rows+1][S-a_x0]
//   region B: pool rows b_r0..b_r1, all S columns              stored [9][b_r1-b_r0+1][S]
// Cells outside both regions read as 0 (the zero-filled pool of the parity contract).
struct CostState {
    void* a;        // nullptr = region empty
    void* b;
    int a_x0, a_rows;
    int b_r0, b_r1;
};

// One plane pass = one thread block sweeping pool rows 1..sweep_rows over all S pool columns.
struct PlaneTask {
    void* plane;            // device pointer to row 0 of the dst plane
    long long pitch;        // elements
    const void* src;        // device pointer to kept row 0 (reference: the rows the BitBlt at GetFrame :361-377 copies)
    long long src_pitch;    // elements between consecutive kept rows. In place: src = plane + offset*pitch, src_pitch = 2*pitch
    int copy_kept;          // 1: the kept rows are not in the dst plane yet - the kernel writes them too
    int no_border;          // 1: the row without a neighbour pair (reference GetFrame :380-391) is not written either: the host
                            // path keeps kept rows and border row on the host and brings home only the interpolated rows
    int width;              // W: samples per row that carry pixels (cost rectangle width)
    int height;             // H: rows of the dst plane
    int offset;             // 0 / 1: first kept row
    int kept_rows;          // n = H/2
    int sweep_rows;         // R >= n-1: pool rows to run the cost recursion over
    int thr_i;              // threshold truncated to the sample type (integer flavours)
    float thr_f;            // fp32 flavour
    int cone;               // dependency cone of everything downstream of this pass (sangnom_plan.h): at pool row r only the
                            // columns < cone - 3r can still reach a picture sample of this or a later pass of the frame;
                            // threads beyond it have nothing left to do. kNoCone: sweep everything (persistent pool).
    int export_cone;        // the same bound for what the NEXT pass of the frame reads of this pass's blurred rows: row r is
                            // handed over only by the threads whose first column is below export_cone - 3r
    CostState in, out;      // cost state from the previous pass / for the next pass of this frame
};
constexpr int kNoCone = 1 << 29;

struct LaunchGeometry {
    int S;                  // pool row length in samples (align32 of the output luma width)
    int Hb;                 // pool rows: (output luma height + 1) >> 1
    unsigned key_mask;      // 0x0FF00FF0 (8-bit kernel): passed as data so that it lives in a register, which lets
                            // (sum & mask) | rank compile to one LOP3; make_geometry() fills it
    int saturate;           // 0: opt=0 arithmetic (narrowing wraps); 1: the SSE2 path's (narrowing saturates) - selects the kernel flavour
    int narrow;             // 1: some plane of the launch is narrower than the pool (subsampled chroma): the 8-bit launcher
                            // brings spare threads and the kernel variant that realigns its warps to the picture edge
};
inline LaunchGeometry make_geometry(int S, int Hb, bool saturate = false, bool narrow = false)
{
    return LaunchGeometry{ S, Hb, 0x0FF00FF0u, saturate ? 1 : 0, narrow ? 1 : 0 };
}

// Widest pool each sample type can run (columns per thread x max threads per block).
int max_pool_width(int sample_bytes);
// Whether a pool row of S samples can be split into whole-thread column segments that fit a block (8-bit pools wider
// than 8192 need S % 64 == 0).
bool pool_width_supported(int sample_bytes, int S);

// Launch one block per task. `tasks_dev` is a device array of ntasks PlaneTask.
// Returns cudaSuccess or the launch error. sample_bytes in {1,2,4}.
cudaError_t launch_plane_tasks(int sample_bytes, const PlaneTask* tasks_dev, int ntasks, LaunchGeometry g,
                               cudaStream_t stream);

// Quarter turns / transposition of whole planes (sangnom_turn.cuh), many planes per launch.
struct TurnPlane {
    const void* src; long long src_pitch;   // W x H samples, pitch in bytes
    void* dst; long long dst_pitch;         // H x W samples
    int width, height;
    // The bytes between the end of a dst row (height samples) and the next multiple of 16 may be overwritten (the
    // tensor-map store of the TMA path clips at 16-byte granularity). The device chain's own planes say yes; planes of
    // a caller who did not say so take the TMA path only when their rows are a whole number of 16-byte pieces.
    int dst_padding_writable;
};
enum TurnKind { kTranspose = 0, kTurnRight = 1, kTurnLeft = 2 };
// Asynchronous on `stream`; the plane table travels in the kernel parameters (64 planes per launch). Returns the
// launch error; *launches (optional) receives the number of kernels launched.
cudaError_t launch_turn_planes(int sample_bytes, const TurnPlane* planes, int nplanes, TurnKind kind, cudaStream_t stream, int* launches = nullptr);

// Name of the kernel variant launch_plane_tasks would use (for logs / profiles).
const char* kernel_variant_name(int sample_bytes, int S);

}  // namespace sn
