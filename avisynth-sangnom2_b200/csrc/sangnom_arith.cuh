// Scalar pixel arithmetic of the reference's opt=0 path for the 16-bit and fp32 flavours
// (/root/reference/src/SangNom2.cpp:36-72, :108-117, :208-249). Integer flavours compute in int32 and
// wrap to the container width where the reference narrows to T; fp32 rounds every operation
// separately in the reference's order (no FMA contraction anywhere: explicit __f*_rn).
#pragma once
#include "sangnom_kernels.h"

#include <cstdint>

namespace sn {

template <typename T> struct Flavour;
template <> struct Flavour<uint8_t>  { using I = int;   static constexpr bool kFloat = false; static constexpr int kMask = 0xFF; };
template <> struct Flavour<uint16_t> { using I = int;   static constexpr bool kFloat = false; static constexpr int kMask = 0xFFFF; };
template <> struct Flavour<float>    { using I = float; static constexpr bool kFloat = true;  static constexpr int kMask = -1; };

// rank of cost buffer i in the reference's tie order: 4 first, then 5,3,6,2,7,1,8,0 (:214-249)
__device__ __forceinline__ constexpr int rank_of(int i)
{
    constexpr int r[kNumCost] = { 8, 6, 4, 2, 0, 1, 3, 5, 7 };
    return r[i];
}

// kSat: the reference's SSE2 flavour (SangNom2_SSE2.cpp:449-517): unsigned lanes, logical shift, pack with saturation,
// i.e. a negative sum becomes the type's maximum and a quotient above it clamps.
template <typename T, bool kSat> __device__ __forceinline__ int tap3_int(int p1, int p2, int p3)
{
    const int s = 4 * p1 + 5 * p2 - p3;
    if constexpr (kSat) return (int)min((unsigned)s >> 3, (unsigned)Flavour<T>::kMask);      // negative -> huge -> max
    else return (s >> 3) & Flavour<T>::kMask;                     // arithmetic shift, then wrap to T (:63-64)
}
__device__ __forceinline__ float tap3_f32(float p1, float p2, float p3)
{
    return __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(p1, 4.0f), __fmul_rn(p2, 5.0f)), p3), 0.125f);   // (:70-71)
}
template <typename T, typename I, bool kSat = false> __device__ __forceinline__ I tap3(I p1, I p2, I p3)
{
    if constexpr (Flavour<T>::kFloat) return tap3_f32(p1, p2, p3);
    else return tap3_int<T, kSat>(p1, p2, p3);
}
__device__ __forceinline__ int absdiff(int a, int b) { return abs(a - b); }
__device__ __forceinline__ float absdiff(float a, float b) { return fabsf(__fsub_rn(a, b)); }
__device__ __forceinline__ int mean2(int a, int b) { return (a + b + 1) >> 1; }
__device__ __forceinline__ float mean2(float a, float b) { return __fmul_rn(__fadd_rn(a, b), 0.5f); }
__device__ __forceinline__ int add2(int a, int b) { return a + b; }
__device__ __forceinline__ float add2(float a, float b) { return __fadd_rn(a, b); }

// The nine raw costs of pixel p. wc/wn: cur/next row windows, pixel p at index p + kHalo.
template <typename T, typename I, int N, int kHalo, bool kSat = false>
__device__ __forceinline__ void raw_costs(const I (&wc)[N], const I (&wn)[N], int p, I (&cost)[kNumCost])
{
    const int q = p + kHalo;
    const I f1 = tap3<T, I, kSat>(wc[q - 1], wc[q], wc[q + 1]);
    const I f2 = tap3<T, I, kSat>(wn[q + 1], wn[q], wn[q - 1]);
    const I b1 = tap3<T, I, kSat>(wc[q + 1], wc[q], wc[q - 1]);
    const I b2 = tap3<T, I, kSat>(wn[q - 1], wn[q], wn[q + 1]);
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

// Interpolated value of pixel p from the winning rank (0 = plain vertical mean).
template <typename T, typename I, int N, int kHalo, bool kSat = false>
__device__ __forceinline__ I interpolate_rank(const I (&wc)[N], const I (&wn)[N], int p, int rank)
{
    const int q = p + kHalo;
    I a = wc[q], b = wn[q];
    // ranks 1 and 2 average two 3-tap values; they differ only in which side neighbour is the first tap
    const bool sg = rank == 1 || rank == 2, rev = rank == 1;
    const I a1 = rev ? wc[q + 1] : wc[q - 1], a3 = rev ? wc[q - 1] : wc[q + 1];
    const I sa = tap3<T, I, kSat>(a1, wc[q], a3);
    const I sb = tap3<T, I, kSat>(rev ? wn[q - 1] : wn[q + 1], wn[q], rev ? wn[q + 1] : wn[q - 1]);
    if (rank == 3) { a = wc[q + 1]; b = wn[q - 1]; }
    if (rank == 4) { a = wc[q - 1]; b = wn[q + 1]; }
    if (rank == 5) { a = wc[q + 2]; b = wn[q - 2]; }
    if (rank == 6) { a = wc[q - 2]; b = wn[q + 2]; }
    if (rank == 7) { a = wc[q + 3]; b = wn[q - 3]; }
    if (rank == 8) { a = wc[q - 3]; b = wn[q + 3]; }
    if (sg) { a = sa; b = sb; }
    return mean2(a, b);
}

}  // namespace sn
