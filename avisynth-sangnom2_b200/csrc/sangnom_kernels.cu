// SangNom2 field interpolation for sm_100a: one fused kernel per sample type.
//
// What it computes (reference /root/reference/src/SangNom2.cpp, opt=0 path):
//   prepareBuffers_c  :74-124   nine direction costs between each pair of kept rows
//   processBuffers_c  :126-159  the in-place, row-RECURSIVE 3x7 cost sum (/16, narrowed to T)
//   finalizePlane_c   :161-257  min-cost direction under the aa threshold + interpolation
// fused so that the nine cost buffers never exist in memory: only the previous blurred row lives
// on chip (registers), because the reference's "blur" is a recursion down the rows
// (B[r] = H7(B[r-1] + P[r] + P[r+1]) / 16), not a box filter.
//
// Decomposition ("row sweep"): one thread block owns one plane pass and sweeps pool rows
// top to bottom; a thread owns COLS adjacent pool columns for all nine costs. Per row the threads
// exchange the 3-column halo of the vertical sums L through shared memory. Parallelism across
// the GPU comes from many (frame, plane) passes in flight, one block each.
//
// The reference's scratch pool is wider/taller than a chroma plane and is never cleared, so what
// the previous plane of the same frame left outside the current plane's rectangle feeds the
// recursion (SURVEY.md finding 2). That state is passed explicitly between passes as CostState
// (sangnom_kernels.h) and every pass runs over the full pool width S, so the result is exact.
#include "sangnom_kernels.h"
#include "sangnom_u8.cuh"

#include <cstdint>
#include <cstdlib>

namespace sn {

namespace {

constexpr int kPad = 4;          // smem row padding on each side (3 clamp columns + alignment)
constexpr int kMaxThreads = 512;

template <typename T> struct Flavour;
template <> struct Flavour<uint8_t>  { using I = int;   static constexpr bool kFloat = false; static constexpr int kMask = 0xFF; };
template <> struct Flavour<uint16_t> { using I = int;   static constexpr bool kFloat = false; static constexpr int kMask = 0xFFFF; };
template <> struct Flavour<float>    { using I = float; static constexpr bool kFloat = true;  static constexpr int kMask = 0; };

// Priority of the nine costs when several equal the minimum (reference :214-249):
// 4 first, then 5,3,6,2,7,1,8,0. kPrio[i] = rank of buffer i; kBufOfRank is the inverse.
__device__ constexpr int kPrio[kNumCost] = { 8, 6, 4, 2, 0, 1, 3, 5, 7 };

// ---- arithmetic primitives (reference :36-72) -------------------------------------------------
template <typename T> __device__ __forceinline__ int tap3_int(int p1, int p2, int p3)
{
    return ((4 * p1 + 5 * p2 - p3) >> 3) & Flavour<T>::kMask;      // arithmetic shift, wrap to T
}
__device__ __forceinline__ float tap3_f32(float p1, float p2, float p3)
{
    // (p1*4 + p2*5 - p3) * 0.125f, every step rounded separately (no FMA contraction)
    return __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(p1, 4.0f), __fmul_rn(p2, 5.0f)), p3), 0.125f);
}
__device__ __forceinline__ int absdiff(int a, int b) { return abs(a - b); }
__device__ __forceinline__ float absdiff(float a, float b) { return fabsf(__fsub_rn(a, b)); }
__device__ __forceinline__ int mean2(int a, int b) { return (a + b + 1) >> 1; }
__device__ __forceinline__ float mean2(float a, float b) { return __fmul_rn(__fadd_rn(a, b), 0.5f); }

template <typename T, typename I>
__device__ __forceinline__ I tap3(I p1, I p2, I p3)
{
    if constexpr (Flavour<T>::kFloat) return tap3_f32(p1, p2, p3);
    else return tap3_int<T>(p1, p2, p3);
}

// ---- cost state between passes ----------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T state_load(const CostState& s, int i, int r, int x, int S)
{
    if (s.b != nullptr && r >= s.b_r0 && r <= s.b_r1) {
        const int nb = s.b_r1 - s.b_r0 + 1;
        return static_cast<const T*>(s.b)[((size_t)i * nb + (r - s.b_r0)) * S + x];
    }
    if (s.a != nullptr && x >= s.a_x0 && r >= 1 && r <= s.a_rows) {
        const int wa = S - s.a_x0;
        return static_cast<const T*>(s.a)[((size_t)i * (s.a_rows + 1) + r) * wa + (x - s.a_x0)];
    }
    return T(0);
}

template <typename T>
__device__ __forceinline__ void state_store(const CostState& s, int i, int r, int x, int S, T v)
{
    if (s.b != nullptr && r >= s.b_r0 && r <= s.b_r1) {
        const int nb = s.b_r1 - s.b_r0 + 1;
        static_cast<T*>(s.b)[((size_t)i * nb + (r - s.b_r0)) * S + x] = v;
    } else if (s.a != nullptr && x >= s.a_x0 && r >= 1 && r <= s.a_rows) {
        const int wa = S - s.a_x0;
        static_cast<T*>(s.a)[((size_t)i * (s.a_rows + 1) + r) * wa + (x - s.a_x0)] = v;
    }
}

// ---- the nine raw costs of one pixel from the two row windows ---------------------------------
// wc/wn hold cur/next pixels x0-3 .. x0+COLS+2; pixel p sits at index p+3.
template <typename T, typename I, int COLS>
__device__ __forceinline__ void raw_costs(const I (&wc)[COLS + 6], const I (&wn)[COLS + 6], int p, I (&cost)[kNumCost])
{
    const int q = p + 3;
    const I f1 = tap3<T, I>(wc[q - 1], wc[q], wc[q + 1]);
    const I f2 = tap3<T, I>(wn[q + 1], wn[q], wn[q - 1]);
    const I b1 = tap3<T, I>(wc[q + 1], wc[q], wc[q - 1]);
    const I b2 = tap3<T, I>(wn[q - 1], wn[q], wn[q + 1]);
    cost[0] = absdiff(wc[q - 3], wn[q + 3]);
    cost[1] = absdiff(wc[q - 2], wn[q + 2]);
    cost[2] = absdiff(wc[q - 1], wn[q + 1]);
    cost[3] = absdiff(f1, f2);
    cost[4] = absdiff(wc[q], wn[q]);
    cost[5] = absdiff(b1, b2);
    cost[6] = absdiff(wc[q + 1], wn[q - 1]);
    cost[7] = absdiff(wc[q + 2], wn[q - 2]);
    cost[8] = absdiff(wc[q + 3], wn[q - 3]);
}

// ---- direction select + interpolation of one pixel (reference :208-249) -----------------------
template <typename T, typename I, int COLS>
__device__ __forceinline__ I interpolate(const I (&wc)[COLS + 6], const I (&wn)[COLS + 6], int p,
                                         const I (&blur)[kNumCost], int thr_i, float thr_f)
{
    const int q = p + 3;
    int rank;   // 0 = plain vertical mean, then 1..8 in the reference's tie order
    if constexpr (Flavour<T>::kFloat) {
        float m = blur[0];
#pragma unroll
        for (int i = 1; i < kNumCost; ++i) m = fminf(m, blur[i]);
        rank = 8;                                   // buffer 0 is the last resort
        if (blur[8] == m) rank = 7;
        if (blur[1] == m) rank = 6;
        if (blur[7] == m) rank = 5;
        if (blur[2] == m) rank = 4;
        if (blur[6] == m) rank = 3;
        if (blur[3] == m) rank = 2;
        if (blur[5] == m) rank = 1;
        if (blur[4] == m || m > thr_f) rank = 0;
    } else {
        // (cost << 4 | rank): one min chain yields the minimum and the tie winner together
        int key = (blur[0] << 4) | kPrio[0];
#pragma unroll
        for (int i = 1; i < kNumCost; ++i) key = min(key, (blur[i] << 4) | kPrio[i]);
        rank = ((key >> 4) > thr_i) ? 0 : (key & 15);
    }
    I a = wc[q], b = wn[q];
    if (rank == 1) { a = tap3<T, I>(wc[q + 1], wc[q], wc[q - 1]); b = tap3<T, I>(wn[q - 1], wn[q], wn[q + 1]); }
    if (rank == 2) { a = tap3<T, I>(wc[q - 1], wc[q], wc[q + 1]); b = tap3<T, I>(wn[q + 1], wn[q], wn[q - 1]); }
    if (rank == 3) { a = wc[q + 1]; b = wn[q - 1]; }
    if (rank == 4) { a = wc[q - 1]; b = wn[q + 1]; }
    if (rank == 5) { a = wc[q + 2]; b = wn[q - 2]; }
    if (rank == 6) { a = wc[q - 2]; b = wn[q + 2]; }
    if (rank == 7) { a = wc[q + 3]; b = wn[q - 3]; }
    if (rank == 8) { a = wc[q - 3]; b = wn[q + 3]; }
    return mean2(a, b);
}

// ---- kernel -------------------------------------------------------------------------------------
template <typename T, int COLS>
__global__ void __launch_bounds__(kMaxThreads)
sangnom_row_sweep(const PlaneTask* __restrict__ tasks, LaunchGeometry g)
{
    using I = typename Flavour<T>::I;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const PlaneTask t = tasks[blockIdx.x];
    const int S = g.S;
    const int LS = S + 2 * kPad;                         // padded smem row length
    I* const Lrow = reinterpret_cast<I*>(smem_raw);        // [9][LS]   vertical sums of the current row
    T* const Pix = reinterpret_cast<T*>(Lrow + kNumCost * LS);   // [3][LS]  ring of kept rows

    const int W = t.width, n = t.kept_rows, R = t.sweep_rows;
    const int x0 = threadIdx.x * COLS;                   // first pool column of this thread
    const bool first_thread = (threadIdx.x == 0);
    const bool last_thread = (x0 + COLS == S);
    T* const plane = static_cast<T*>(t.plane);
    const long long pitch = t.pitch;

    auto kept_row = [&](int j) -> T* { return plane + (long long)(t.offset + 2 * j) * pitch; };

    // Publish this thread's part of kept row j into ring slot j%3, with the 3 clamp pixels on each
    // side of the picture (reference loadPixel :25-34 replicates the edge pixel).
    T pre[COLS];
    auto fetch_row = [&](int j) {
        const T* src = kept_row(j);
#pragma unroll
        for (int c = 0; c < COLS; ++c) pre[c] = (x0 + c < W) ? src[x0 + c] : T(0);
    };
    auto publish_row = [&](int j) {
        T* dst = Pix + (j % 3) * LS + kPad;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            if (x0 + c < W) dst[x0 + c] = pre[c];
            if (x0 + c == W - 1) { dst[W] = pre[c]; dst[W + 1] = pre[c]; dst[W + 2] = pre[c]; }   // right clamp
        }
        if (first_thread) { dst[-1] = pre[0]; dst[-2] = pre[0]; dst[-3] = pre[0]; }                // left clamp
    };
    auto load_window = [&](int j, I (&w)[COLS + 6]) {
        const T* src = Pix + (j % 3) * LS + kPad + x0 - 3;
#pragma unroll
        for (int c = 0; c < COLS + 6; ++c) w[c] = static_cast<I>(src[c]);
    };

    // ---- border row that has no neighbour pair (reference GetFrame :380-391) ----
    {
        const T* from = (t.offset == 0) ? plane + (long long)(t.height - 2) * pitch : plane + pitch;
        T* to = (t.offset == 0) ? plane + (long long)(t.height - 1) * pitch : plane;
#pragma unroll
        for (int c = 0; c < COLS; ++c) if (x0 + c < W) to[x0 + c] = from[x0 + c];
    }

    // ---- prologue: rows K0, K1 (K2 prefetched), M = B[0] + P[1] with B[0] = 0 ----
    for (int j = 0; j < 2 && j < n; ++j) { fetch_row(j); publish_row(j); }
    __syncthreads();

    I wa[COLS + 6], wb[COLS + 6], wc[COLS + 6];            // windows of K[r-1], K[r], K[r+1]
    I M[kNumCost][COLS];                                  // B[r-1] + P[r]
    load_window(0, wa);
    if (n >= 2) load_window(1, wb);
    {
        const bool pair = (1 <= n - 1);
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            I cost[kNumCost];
            if (pair && x0 + c < W) {
                raw_costs<T, I, COLS>(wa, wb, c, cost);
            } else {
#pragma unroll
                for (int i = 0; i < kNumCost; ++i) cost[i] = static_cast<I>(state_load<T>(t.in, i, 1, x0 + c, S));
            }
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) M[i][c] = cost[i];
        }
    }
    if (2 < n) { fetch_row(2); publish_row(2); }
    __syncthreads();

    // ---- sweep ----
    for (int r = 1; r <= R; ++r) {
        // wa = K[r-1], wb = K[r]
        const bool prefetch = (r + 2 <= n - 1);
        if (prefetch) fetch_row(r + 2);

        const bool pair = (r + 1 <= n - 1);                // P[r+1] comes from pixels K[r], K[r+1]
        if (pair) load_window(r + 1, wc);

        I Pn[kNumCost][COLS];
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            I cost[kNumCost];
            if (pair && x0 + c < W) {
                raw_costs<T, I, COLS>(wb, wc, c, cost);
            } else {
#pragma unroll
                for (int i = 0; i < kNumCost; ++i) cost[i] = static_cast<I>(state_load<T>(t.in, i, r + 1, x0 + c, S));
            }
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) Pn[i][c] = cost[i];
        }

        // vertical sums L = (B[r-1] + P[r]) + P[r+1]; publish for the neighbours.
        // M then keeps P[r+1] until the blurred row is known.
#pragma unroll
        for (int i = 0; i < kNumCost; ++i) {
            I* row = Lrow + i * LS + kPad;
            I first = I(0), last = I(0);
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                I v;
                if constexpr (Flavour<T>::kFloat) v = __fadd_rn(M[i][c], Pn[i][c]); else v = M[i][c] + Pn[i][c];
                row[x0 + c] = v;
                M[i][c] = Pn[i][c];
                if (c == 0) first = v;
                if (c == COLS - 1) last = v;
            }
            if (first_thread) { row[-1] = first; row[-2] = first; row[-3] = first; }     // clamp at column 0
            if (last_thread) { row[S] = last; row[S + 1] = last; row[S + 2] = last; }    // clamp at column S-1
        }
        __syncthreads();

        // horizontal 7-tap, /16, narrow to T  (reference :144-152)
        I B[kNumCost][COLS];
#pragma unroll
        for (int i = 0; i < kNumCost; ++i) {
            const I* row = Lrow + i * LS + kPad + x0 - 3;
            I L[COLS + 6];
#pragma unroll
            for (int c = 0; c < COLS + 6; ++c) L[c] = row[c];
            if constexpr (Flavour<T>::kFloat) {
#pragma unroll
                for (int c = 0; c < COLS; ++c) {
                    float s = __fadd_rn(L[c], L[c + 1]);
                    s = __fadd_rn(s, L[c + 2]);
                    s = __fadd_rn(s, L[c + 3]);
                    s = __fadd_rn(s, L[c + 4]);
                    s = __fadd_rn(s, L[c + 5]);
                    s = __fadd_rn(s, L[c + 6]);
                    B[i][c] = __fmul_rn(s, 0.0625f);
                }
            } else {
                int s = L[0] + L[1] + L[2] + L[3] + L[4] + L[5] + L[6];
                B[i][0] = (s >> 4) & Flavour<T>::kMask;
#pragma unroll
                for (int c = 1; c < COLS; ++c) {
                    s += L[c + 6] - L[c - 1];
                    B[i][c] = (s >> 4) & Flavour<T>::kMask;
                }
            }
        }

        // interpolate picture row between K[r-1] and K[r]
        if (r <= n - 1) {
            T* out = plane + (long long)(t.offset + 2 * (r - 1) + 1) * pitch;
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                if (x0 + c < W) {
                    I b9[kNumCost];
#pragma unroll
                    for (int i = 0; i < kNumCost; ++i) b9[i] = B[i][c];
                    out[x0 + c] = static_cast<T>(interpolate<T, I, COLS>(wa, wb, c, b9, t.thr_i, t.thr_f));
                }
            }
        }

        // hand the blurred row to the next pass of this frame where it will look outside its rectangle
        if (t.out.a != nullptr || t.out.b != nullptr) {
#pragma unroll
            for (int i = 0; i < kNumCost; ++i)
#pragma unroll
                for (int c = 0; c < COLS; ++c) state_store<T>(t.out, i, r, x0 + c, S, static_cast<T>(B[i][c]));
        }

        // next row's running term and windows
#pragma unroll
        for (int i = 0; i < kNumCost; ++i)
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                if constexpr (Flavour<T>::kFloat) M[i][c] = __fadd_rn(B[i][c], M[i][c]); else M[i][c] = B[i][c] + M[i][c];
            }
#pragma unroll
        for (int c = 0; c < COLS + 6; ++c) { wa[c] = wb[c]; wb[c] = wc[c]; }

        if (prefetch) publish_row(r + 2);
        __syncthreads();
    }
}

template <typename T, int COLS>
cudaError_t launch_variant(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream)
{
    using I = typename Flavour<T>::I;
    const int threads = g.S / COLS;
    const size_t smem = (size_t)(g.S + 2 * kPad) * (kNumCost * sizeof(I) + 3 * sizeof(T));
    static size_t configured[64] = {};          // per device: largest dynamic smem opted into so far
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || smem > configured[dev]) {
        e = cudaFuncSetAttribute(sangnom_row_sweep<T, COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = smem;
    }
    sangnom_row_sweep<T, COLS><<<ntasks, threads, smem, stream>>>(tasks, g);
    return cudaGetLastError();
}

template <int kMaxThreads, int kMinBlocks>
cudaError_t launch_u8(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream)
{
    const size_t smem = u8k::smem_bytes(g.S);
    static size_t configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || smem > configured[dev]) {
        e = cudaFuncSetAttribute(u8k::sangnom_u8_row_sweep<kMaxThreads, kMinBlocks>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = smem;
    }
    u8k::sangnom_u8_row_sweep<kMaxThreads, kMinBlocks><<<ntasks, g.S / u8k::kCols, smem, stream>>>(tasks, g);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_typed(const PlaneTask* tasks, int ntasks, LaunchGeometry g, cudaStream_t stream)
{
    if (g.S <= 4 * kMaxThreads) return launch_variant<T, 4>(tasks, ntasks, g, stream);
    return launch_variant<T, 8>(tasks, ntasks, g, stream);
}

}  // namespace

int max_pool_width(int sample_bytes)
{
    (void)sample_bytes;
    return 8 * kMaxThreads;
}

const char* kernel_variant_name(int sample_bytes, int S)
{
    const bool wide = S > 4 * kMaxThreads;
    switch (sample_bytes) {
        case 1: return S <= 2048 ? "sangnom_u8_row_sweep<256,3>" : "sangnom_u8_row_sweep<512,1>";
        case 2: return wide ? "sangnom_row_sweep<u16,8>" : "sangnom_row_sweep<u16,4>";
        default: return wide ? "sangnom_row_sweep<f32,8>" : "sangnom_row_sweep<f32,4>";
    }
}

cudaError_t launch_plane_tasks(int sample_bytes, const PlaneTask* tasks_dev, int ntasks, LaunchGeometry g, cudaStream_t stream)
{
    if (ntasks <= 0) return cudaSuccess;
    if (g.S % 32 != 0 || g.S > max_pool_width(sample_bytes)) return cudaErrorInvalidValue;
    switch (sample_bytes) {
        case 1: {
            static const int occ = [] { const char* v = getenv("SANGNOM_U8_OCC"); return v ? atoi(v) : 2; }();   // tuning knob
            if (occ == 0) return launch_typed<uint8_t>(tasks_dev, ntasks, g, stream);                           // generic kernel
            if (g.S <= 256 * u8k::kCols) {
                if (occ >= 3) return launch_u8<256, 3>(tasks_dev, ntasks, g, stream);
                return launch_u8<256, 2>(tasks_dev, ntasks, g, stream);
            }
            return launch_u8<512, 1>(tasks_dev, ntasks, g, stream);
        }
        case 2: return launch_typed<uint16_t>(tasks_dev, ntasks, g, stream);
        case 4: return launch_typed<float>(tasks_dev, ntasks, g, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sn
