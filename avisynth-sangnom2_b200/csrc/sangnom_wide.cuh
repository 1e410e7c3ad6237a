// 16-bit and fp32 flavours of the fused row sweep (one pixel per 32-bit lane), with optional
// column split of a plane over the blocks of a thread-block cluster.
//
// A thread owns 4 adjacent pool columns for all nine costs. State carried down the rows:
// M = B[r-1] + P[r] (36 registers). Per pool row: raw costs P[r+1] from the two kept rows
// (read straight from global/L1: each row is touched by the same thread in three consecutive
// iterations), L = M + P[r+1] into a double-buffered shared row, ONE barrier, then per cost the
// 7-tap sum, /16, narrow to T, min-key update and M = B + P[r+1]; finally the interpolated row.
// When the plane is split, the edge threads push their three edge L values into the neighbour
// block's shared row (DSMEM) and the barrier is the cluster barrier.
//
// Reference semantics: /root/reference/src/SangNom2.cpp :74-124, :126-159, :161-257.
#pragma once
#include "sangnom_arith.cuh"
#include "sangnom_cluster.cuh"

#ifndef SN_DYNAMIC_SMEM
#define SN_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace sn {
namespace wide {

constexpr int kCols = 4;                 // pool columns per thread
constexpr int kHalo = 4;                 // window / shared-row halo in elements (3 are used; 4 keeps 16-byte alignment)
constexpr int kWin = kCols + 2 * kHalo;  // 12

__device__ __forceinline__ void prefetch_l1(const void* p)
{
#ifndef SN_HOST_EMULATION
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// ---- 4-element vector access -----------------------------------------------------------------
__device__ __forceinline__ void load4(const uint16_t* p, int (&v)[4])
{
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    v[0] = (int)(r.x & 0xFFFFu); v[1] = (int)(r.x >> 16); v[2] = (int)(r.y & 0xFFFFu); v[3] = (int)(r.y >> 16);
}
__device__ __forceinline__ void load4(const float* p, float (&v)[4])
{
    const float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
__device__ __forceinline__ void store4(uint16_t* p, const int (&v)[4])
{
    *reinterpret_cast<uint2*>(p) = make_uint2(((uint32_t)v[0] & 0xFFFFu) | ((uint32_t)v[1] << 16), ((uint32_t)v[2] & 0xFFFFu) | ((uint32_t)v[3] << 16));
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4])
{
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// Window of elements x0-4 .. x0+7 of a picture row, edges replicated (reference loadPixel :25-34).
template <typename T, typename I>
__device__ __forceinline__ void load_window(const T* __restrict__ row, int x0, int W, bool vec, I (&w)[kWin])
{
#pragma unroll
    for (int e = 0; e < kWin; ++e) w[e] = I(0);
    if (x0 >= W) return;
    if (vec) {
        I v[4];
        if (x0 > 0) { load4(row + x0 - 4, v); w[0] = v[0]; w[1] = v[1]; w[2] = v[2]; w[3] = v[3]; }
        load4(row + x0, v); w[4] = v[0]; w[5] = v[1]; w[6] = v[2]; w[7] = v[3];
        if (x0 + 4 < W) { load4(row + x0 + 4, v); w[8] = v[0]; w[9] = v[1]; w[10] = v[2]; w[11] = v[3]; }
    } else {
#pragma unroll
        for (int e = 0; e < kWin; ++e) { const int x = x0 - kHalo + e; if (x >= 0 && x < W) w[e] = (I)row[x]; }
    }
    if (x0 == 0) { w[0] = w[4]; w[1] = w[4]; w[2] = w[4]; w[3] = w[4]; }
    const int last = W - 1 - (x0 - kHalo);      // window index of the last picture column (>= kHalo)
#pragma unroll
    for (int e = kHalo + 1; e < kWin; ++e) if (e > last) w[e] = w[e - 1];
}

template <typename T, int kMaxThreads, int kMinBlocks, bool kClustered, bool kSat = false>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks)
sangnom_wide_row_sweep(const PlaneTask* __restrict__ tasks, LaunchGeometry g, int seg_cols)
{
    using I = typename Flavour<T>::I;
    SN_DYNAMIC_SMEM(smem_raw);

    const unsigned G = kClustered ? cl::size() : 1u;      // kClustered = false: one block per plane, no cluster code at all
    const unsigned crank = kClustered ? cl::rank() : 0u;
    const PlaneTask t = tasks[blockIdx.x / G];
    const int S = g.S;
    const int LS = seg_cols + 2 * kHalo;                         // elements per shared L row of this segment
    I* const Lbase = reinterpret_cast<I*>(smem_raw);             // [2][9][LS]

    const int W = t.width, n = t.kept_rows, R = t.sweep_rows;
    const int lx = threadIdx.x * kCols;                          // column inside the segment
    const int x0 = (int)crank * seg_cols + lx;                   // pool column
    const bool plane_first = x0 == 0, plane_last = x0 + kCols == S;
    const bool seg_first = lx == 0, seg_last = lx + kCols == seg_cols;
    T* const plane = static_cast<T*>(t.plane);
    const T* const src = static_cast<const T*>(t.src);
    const long long pitch = t.pitch, src_pitch = t.src_pitch;
    const long long wpad = ((long long)W * (long long)sizeof(T) + 15) & ~15LL;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(src_pitch * (long long)sizeof(T))) & 15) == 0 &&
                     src_pitch * (long long)sizeof(T) >= wpad;                       // aligned vector loads of kept rows
    const bool vec_out = ((reinterpret_cast<uintptr_t>(plane) | (uintptr_t)(pitch * (long long)sizeof(T))) & 15) == 0 &&
                         pitch * (long long)sizeof(T) >= wpad;                       // aligned vector stores
    const int npix = min(max(W - x0, 0), kCols);                 // how many of my columns carry pixels

    auto kept_row = [&](int j) -> const T* { return src + (long long)j * src_pitch; };
    auto store_px = [&](T* row, const I (&v)[4]) {
        if (npix == kCols && vec_out) { store4(row + x0, v); return; }
#pragma unroll
        for (int c = 0; c < kCols; ++c) if (c < npix) row[x0 + c] = (T)v[c];
    };

    // ---- border row without a neighbour pair (reference GetFrame :380-391) ----
    if (npix > 0) {
        T* to = t.offset == 0 ? plane + (long long)(t.height - 1) * pitch : plane;
        I w[kWin];
        load_window<T, I>(kept_row(t.offset == 0 ? n - 1 : 0), x0, W, vec, w);
        const I own[4] = { w[4], w[5], w[6], w[7] };
        store_px(to, own);
        if (t.copy_kept) {                       // last kept row; rows 0..n-2 are written as the sweep passes them
            if (t.offset != 0) load_window<T, I>(kept_row(n - 1), x0, W, vec, w);
            const I lastrow[4] = { w[4], w[5], w[6], w[7] };
            store_px(plane + (long long)(t.offset + 2 * (n - 1)) * pitch, lastrow);
        }
    }

    // Cost state of the previous pass at pool row `row` for my 4 columns: base pointer and per-buffer stride
    // (elements), or nullptr when the cells read as the pool's zero.
    auto state_row = [&](const CostState& s, int row, size_t& stride) -> T* {
        if (s.b != nullptr && row >= s.b_r0 && row <= s.b_r1) {
            stride = (size_t)(s.b_r1 - s.b_r0 + 1) * S;
            return static_cast<T*>(s.b) + (size_t)(row - s.b_r0) * S + x0;
        }
        if (s.a != nullptr && x0 >= s.a_x0 && row >= 1 && row <= s.a_rows) {
            const int wa = S - s.a_x0;
            stride = (size_t)(s.a_rows + 1) * wa;
            return static_cast<T*>(s.a) + (size_t)row * wa + (x0 - s.a_x0);
        }
        stride = 0;
        return nullptr;
    };

    // Raw cost row `row` of the pool into P: pixels where I have them and the pair exists (windows wc = K[row-1],
    // wn = K[row]), else the handed-over state.
    auto cost_row = [&](int row, const I (&wc)[kWin], const I (&wn)[kWin], I (&P)[kNumCost][kCols]) {
        const bool pair = row <= n - 1;          // pool row j+1 holds the costs of the pair (K[j], K[j+1])
        const bool pixels = pair && npix > 0;
        if (!(pixels && npix == kCols)) {
            size_t stride;
            const T* st = state_row(t.in, row, stride);
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                if (st != nullptr) load4(st + i * stride, P[i]);
                else { P[i][0] = P[i][1] = P[i][2] = P[i][3] = I(0); }
            }
        }
        if (pixels) {
#pragma unroll
            for (int c = 0; c < kCols; ++c) {
                I cost[kNumCost];
                raw_costs<T, I, kWin, kHalo, kSat>(wc, wn, c, cost);
                if (c < npix) {
#pragma unroll
                    for (int i = 0; i < kNumCost; ++i) P[i][c] = cost[i];
                }
            }
        }
    };

    // Rolling register windows of three kept rows: wa = K[r-1], wb = K[r], wc = K[r+1]; each kept row is read from
    // global memory once. M = B[r-1] + P[r] (B[0] = 0) is the only cost state carried down the rows.
    I wa[kWin], wb[kWin], wc[kWin];
    I M[kNumCost][kCols];
#pragma unroll
    for (int e = 0; e < kWin; ++e) { wa[e] = I(0); wb[e] = I(0); wc[e] = I(0); }
    if (npix > 0) {
        load_window<T, I>(kept_row(0), x0, W, vec, wb);
        if (n > 1) load_window<T, I>(kept_row(1), x0, W, vec, wc);
    }
    cost_row(1, wb, wc, M);
    // after this the loop invariant holds at r = 1: wa = K[0], wb = K[1]
#pragma unroll
    for (int e = 0; e < kWin; ++e) { wa[e] = wb[e]; wb[e] = wc[e]; }

    const int tkey = (int)min((long long)t.thr_i + 1, 0x7FFFFFFLL) << 4;      // (thr+1) << 4: "every cost above the threshold"
    const bool exporting = t.out.a != nullptr || t.out.b != nullptr;

    for (int r = 1; r <= R; ++r) {
        // pull the kept row two iterations ahead towards L1 (no register cost)
        if (r + 3 <= n - 1 && npix > 0) prefetch_l1(kept_row(r + 3) + x0);

        // ---- P[r+1]; L = M + P[r+1] into the shared row (and the neighbours' pads); M keeps P[r+1] ----
        if (r + 1 <= n - 1 && npix > 0) load_window<T, I>(kept_row(r + 1), x0, W, vec, wc);
        I* const Lrow = Lbase + (size_t)(r & 1) * kNumCost * LS + kHalo;
        {
            I P[kNumCost][kCols];
            cost_row(r + 1, wb, wc, P);
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                I* row = Lrow + i * LS;
                I L[4];
#pragma unroll
                for (int c = 0; c < kCols; ++c) { L[c] = add2(M[i][c], P[i][c]); M[i][c] = P[i][c]; }
                if constexpr (Flavour<T>::kFloat) *reinterpret_cast<float4*>(row + lx) = make_float4(L[0], L[1], L[2], L[3]);
                else *reinterpret_cast<uint4*>(row + lx) = make_uint4((uint32_t)L[0], (uint32_t)L[1], (uint32_t)L[2], (uint32_t)L[3]);
                if (seg_first) {
                    if (plane_first) { row[-1] = L[0]; row[-2] = L[0]; row[-3] = L[0]; }                         // clamp at column 0
                    else { cl::store_remote(row + seg_cols, crank - 1, L[0]); cl::store_remote(row + seg_cols + 1, crank - 1, L[1]);
                           cl::store_remote(row + seg_cols + 2, crank - 1, L[2]); }
                }
                if (seg_last) {
                    if (plane_last) { row[seg_cols] = L[3]; row[seg_cols + 1] = L[3]; row[seg_cols + 2] = L[3]; } // clamp at column S-1
                    else { cl::store_remote(row - 3, crank + 1, L[1]); cl::store_remote(row - 2, crank + 1, L[2]);
                           cl::store_remote(row - 1, crank + 1, L[3]); }
                }
            }
        }
        if constexpr (kClustered) cl::sync_all(); else __syncthreads();

        // ---- B[r] per cost buffer, min key, M = B[r] + P[r+1], hand-over ----
        size_t out_stride = 0;
        T* const out_ptr = exporting ? state_row(t.out, r, out_stride) : nullptr;
        int kmin[kCols];        // integer flavours: min over (cost << 4 | rank) keys, threshold folded in
        float fmin[kCols];      // fp32: running minimum and the rank that first reached it
        int frank[kCols];
#pragma unroll
        for (int c = 0; c < kCols; ++c) { kmin[c] = tkey; fmin[c] = 0.f; frank[c] = 0; }
        // Buffers are visited in the reference's tie order (4,5,3,6,2,7,1,8,0), so for fp32 a strict
        // "<" update leaves the first of several equal minima as the winner (:214-249).
#pragma unroll
        for (int k = 0; k < kNumCost; ++k) {
            constexpr int order[kNumCost] = { 4, 5, 3, 6, 2, 7, 1, 8, 0 };
            const int i = order[k];
            const I* row = Lrow + i * LS + lx;
            I Lw[kWin];
            if constexpr (Flavour<T>::kFloat) {
                const float4 a = *reinterpret_cast<const float4*>(row - 4), b = *reinterpret_cast<const float4*>(row), c4 = *reinterpret_cast<const float4*>(row + 4);
                Lw[0] = a.x; Lw[1] = a.y; Lw[2] = a.z; Lw[3] = a.w; Lw[4] = b.x; Lw[5] = b.y; Lw[6] = b.z; Lw[7] = b.w; Lw[8] = c4.x; Lw[9] = c4.y; Lw[10] = c4.z; Lw[11] = c4.w;
            } else {
                const uint4 a = *reinterpret_cast<const uint4*>(row - 4), b = *reinterpret_cast<const uint4*>(row), c4 = *reinterpret_cast<const uint4*>(row + 4);
                Lw[0] = (int)a.x; Lw[1] = (int)a.y; Lw[2] = (int)a.z; Lw[3] = (int)a.w; Lw[4] = (int)b.x; Lw[5] = (int)b.y; Lw[6] = (int)b.z; Lw[7] = (int)b.w;
                Lw[8] = (int)c4.x; Lw[9] = (int)c4.y; Lw[10] = (int)c4.z; Lw[11] = (int)c4.w;
            }
            I B4[kCols];
            // pixel c sits at window index c+4; its seven taps are indices c+1 .. c+7
            if constexpr (Flavour<T>::kFloat) {
#pragma unroll
                for (int c = 0; c < kCols; ++c) {
                    float s = __fadd_rn(Lw[c + 1], Lw[c + 2]);                     // ((((((m3+m2)+m1)+c)+p1)+p2)+p3) / 16  (:152)
                    s = __fadd_rn(s, Lw[c + 3]); s = __fadd_rn(s, Lw[c + 4]); s = __fadd_rn(s, Lw[c + 5]);
                    s = __fadd_rn(s, Lw[c + 6]); s = __fadd_rn(s, Lw[c + 7]);
                    B4[c] = __fmul_rn(s, 0.0625f);
                    if (k == 0 || B4[c] < fmin[c]) { fmin[c] = B4[c]; frank[c] = k; }
                }
            } else {
                int s = Lw[1] + Lw[2] + Lw[3] + Lw[4] + Lw[5] + Lw[6] + Lw[7];
#pragma unroll
                for (int c = 0; c < kCols; ++c) {
                    if (c > 0) s += Lw[c + 7] - Lw[c];
                    // narrowT(s / 16) << 4 | rank: wrapped (:152), or clamped for the SSE2 flavour (SangNom2_SSE2.cpp:807)
                    const unsigned kept = kSat ? min((unsigned)s, ((unsigned)Flavour<T>::kMask << 4) | 15u) & ~15u : (unsigned)s & ((unsigned)Flavour<T>::kMask << 4);
                    const int key = (int)(kept | (unsigned)k);
                    B4[c] = key >> 4;
                    kmin[c] = min(kmin[c], key);
                }
            }
#pragma unroll
            for (int c = 0; c < kCols; ++c) M[i][c] = add2(B4[c], M[i][c]);
            if (out_ptr != nullptr) store4(out_ptr + i * out_stride, B4);
        }

        // ---- interpolate the picture row between K[r-1] and K[r] ----
        if (r <= n - 1 && npix > 0) {
            I px[kCols];
#pragma unroll
            for (int c = 0; c < kCols; ++c) {
                int rank;
                if constexpr (Flavour<T>::kFloat) rank = fmin[c] > t.thr_f ? 0 : frank[c];
                else rank = kmin[c] & 15;
                px[c] = interpolate_rank<T, I, kWin, kHalo, kSat>(wa, wb, c, rank);
            }
            store_px(plane + (long long)(t.offset + 2 * (r - 1) + 1) * pitch, px);
            if (t.copy_kept) {
                const I keptrow[4] = { wa[4], wa[5], wa[6], wa[7] };
                store_px(plane + (long long)(t.offset + 2 * (r - 1)) * pitch, keptrow);
            }
        }
#pragma unroll
        for (int e = 0; e < kWin; ++e) { wa[e] = wb[e]; wb[e] = wc[e]; }
    }
}

template <typename T> inline size_t smem_bytes(int seg_cols) { return (size_t)2 * kNumCost * (seg_cols + 2 * kHalo) * sizeof(typename Flavour<T>::I); }

}  // namespace wide
}  // namespace sn
