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

#include <type_traits>

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

    // window of kept row j: interior threads of an aligned plane take three vector loads, nothing else
    const bool interior = vec && x0 >= kHalo && x0 + kCols + kHalo <= W;
    auto window = [&](int j, I (&w)[kWin]) {
        const T* row = kept_row(j);
        if (interior) {
            I v[4];
            load4(row + x0 - 4, v); w[0] = v[0]; w[1] = v[1]; w[2] = v[2]; w[3] = v[3];
            load4(row + x0, v); w[4] = v[0]; w[5] = v[1]; w[6] = v[2]; w[7] = v[3];
            load4(row + x0 + 4, v); w[8] = v[0]; w[9] = v[1]; w[10] = v[2]; w[11] = v[3];
        } else {
            load_window<T, I>(row, x0, W, vec, w);
        }
    };

    // Raw cost row `row` of the pool into P (windows wc = K[row-1], wn = K[row]).
    // kFull: all my columns carry pixels. kPair: the pair (K[row-1], K[row]) exists. Where there are no pixel costs
    // the handed-over state of the previous pass (or the pool's zero) stands in.
    auto cost_row = [&](auto full, auto pairrow, int row, const I (&wc)[kWin], const I (&wn)[kWin], I (&P)[kNumCost][kCols]) {
        constexpr bool kFull = decltype(full)::value, kPair = decltype(pairrow)::value;
        if constexpr (!(kFull && kPair)) {
            size_t stride;
            const T* st = state_row(t.in, row, stride);
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                if (st != nullptr) load4(st + i * stride, P[i]);
                else { P[i][0] = P[i][1] = P[i][2] = P[i][3] = I(0); }
            }
        }
        if constexpr (kPair) {
            if (kFull || npix > 0) {
#pragma unroll
                for (int c = 0; c < kCols; ++c) {
                    I cost[kNumCost];
                    raw_costs<T, I, kWin, kHalo, kSat>(wc, wn, c, cost);
                    if (kFull || c < npix) {
#pragma unroll
                        for (int i = 0; i < kNumCost; ++i) P[i][c] = cost[i];
                    }
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
        window(0, wb);
        if (n > 1) window(1, wc);
    }
    if (n > 1) cost_row(std::false_type{}, std::true_type{}, 1, wb, wc, M);
    else cost_row(std::false_type{}, std::false_type{}, 1, wb, wc, M);
    // after this the loop invariant holds at r = 1: wa = K[0], wb = K[1]
#pragma unroll
    for (int e = 0; e < kWin; ++e) { wa[e] = wb[e]; wb[e] = wc[e]; }

    // all blocks of a cluster run before the first DSMEM store
    if constexpr (kClustered) cl::sync_all();

    const int tkey = (int)min((long long)t.thr_i + 1, 0x7FFFFFFLL) << 4;      // (thr+1) << 4: "every cost above the threshold"
    const bool exporting = t.out.a != nullptr || t.out.b != nullptr;

    // One pool row. kFull: every thread of the warp owns 4 pixel columns. kPair: pool row r+1 is a pair row
    // (r + 1 <= n - 1). kExport: some thread of the warp hands this row's blurred costs to the next pass.
    auto row_step = [&](auto full, auto pairrow, auto exportrow, int r) {
        constexpr bool kFull = decltype(full)::value, kPair = decltype(pairrow)::value, kExport = decltype(exportrow)::value;
        const bool pixels = kFull || npix > 0;
        // pull the kept row two iterations ahead towards L1 (no register cost)
        if (kPair && pixels && r + 3 <= n - 1) prefetch_l1(kept_row(r + 3) + x0);

        // ---- P[r+1]; L = M + P[r+1] into the shared row; M keeps P[r+1] ----
        if (kPair && pixels) window(r + 1, wc);
        I* const Lrow = Lbase + (size_t)(r & 1) * kNumCost * LS + kHalo;
        {
            I P[kNumCost][kCols];
            cost_row(full, pairrow, r + 1, wb, wc, P);
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                I L[4];
#pragma unroll
                for (int c = 0; c < kCols; ++c) { L[c] = add2(M[i][c], P[i][c]); M[i][c] = P[i][c]; }
                I* row = Lrow + i * LS;
                if constexpr (Flavour<T>::kFloat) *reinterpret_cast<float4*>(row + lx) = make_float4(L[0], L[1], L[2], L[3]);
                else *reinterpret_cast<uint4*>(row + lx) = make_uint4((uint32_t)L[0], (uint32_t)L[1], (uint32_t)L[2], (uint32_t)L[3]);
            }
        }
        // the segment's edge threads fill the 3-column pads: clamp at the pool's ends (:144-152 clamps at the pool
        // stride), the neighbour block's shared row otherwise. Their own L values are re-read from the row just written.
        if (seg_first | seg_last) {
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                I* row = Lrow + i * LS;
                if (seg_first) {
                    const I l0 = row[lx], l1 = row[lx + 1], l2 = row[lx + 2];
                    if (plane_first) { row[-1] = l0; row[-2] = l0; row[-3] = l0; }
                    else { cl::store_remote(row + seg_cols, crank - 1, l0); cl::store_remote(row + seg_cols + 1, crank - 1, l1);
                           cl::store_remote(row + seg_cols + 2, crank - 1, l2); }
                }
                if (seg_last) {
                    const I l1 = row[lx + 1], l2 = row[lx + 2], l3 = row[lx + 3];
                    if (plane_last) { row[seg_cols] = l3; row[seg_cols + 1] = l3; row[seg_cols + 2] = l3; }
                    else { cl::store_remote(row - 3, crank + 1, l1); cl::store_remote(row - 2, crank + 1, l2);
                           cl::store_remote(row - 1, crank + 1, l3); }
                }
            }
        }
        if constexpr (kClustered) cl::sync_all(); else __syncthreads();

        // ---- B[r] per cost buffer, min key, M = B[r] + P[r+1], hand-over ----
        size_t out_stride = 0;
        T* out_ptr = nullptr;
        if constexpr (kExport) out_ptr = state_row(t.out, r, out_stride);
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
            if constexpr (kExport) { if (out_ptr != nullptr) store4(out_ptr + i * out_stride, B4); }
        }

        // ---- interpolate the picture row between K[r-1] and K[r] ----
        if ((kPair || r <= n - 1) && pixels) {
            I px[kCols];
#pragma unroll
            for (int c = 0; c < kCols; ++c) {
                int rank;
                if constexpr (Flavour<T>::kFloat) rank = fmin[c] > t.thr_f ? 0 : frank[c];
                else rank = kmin[c] & 15;
                px[c] = interpolate_rank<T, I, kWin, kHalo, kSat>(wa, wb, c, rank);
            }
            T* const orow = plane + (long long)(t.offset + 2 * (r - 1) + 1) * pitch;
            if (kFull && vec_out) store4(orow + x0, px); else store_px(orow, px);
            if (t.copy_kept) {
                const I keptrow[4] = { wa[4], wa[5], wa[6], wa[7] };
                if (kFull && vec_out) store4(orow - pitch + x0, keptrow); else store_px(orow - pitch, keptrow);
            }
        }
#pragma unroll
        for (int e = 0; e < kWin; ++e) { wa[e] = wb[e]; wb[e] = wc[e]; }
    };
    // does any thread of my warp export pool row r? (warp-uniform, so the variants of a row keep warps whole)
    auto export_row = [&](int r) -> bool {
        if (!exporting) return false;
        size_t stride;
        const bool mine = state_row(t.out, r, stride) != nullptr;
#ifdef SN_HOST_EMULATION
        return mine;
#else
        return __any_sync(0xFFFFFFFFu, mine);
#endif
    };
    auto sweep = [&](auto full) {
        int r = 1;
        for (; r <= n - 2 && r <= R; ++r) {                                  // rows whose lower neighbour row is a pair row
            if (export_row(r)) row_step(full, std::true_type{}, std::true_type{}, r);
            else row_step(full, std::true_type{}, std::false_type{}, r);
        }
        for (; r <= R; ++r) {                                               // the last picture row and rows swept for the next pass only
            if (export_row(r)) row_step(full, std::false_type{}, std::true_type{}, r);
            else row_step(full, std::false_type{}, std::false_type{}, r);
        }
    };
    // warp-uniform choice, so that a warp never splits over the copies of the row barrier
#ifdef SN_HOST_EMULATION
    const bool warp_full = npix == kCols;
#else
    const bool warp_full = __all_sync(0xFFFFFFFFu, npix == kCols);
#endif
    if (warp_full) sweep(std::true_type{}); else sweep(std::false_type{});
}

template <typename T> inline size_t smem_bytes(int seg_cols) { return (size_t)2 * kNumCost * (seg_cols + 2 * kHalo) * sizeof(typename Flavour<T>::I); }

}  // namespace wide
}  // namespace sn
