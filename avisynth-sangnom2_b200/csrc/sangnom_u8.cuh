// 8-bit flavour of the fused row sweep, written for the ALU-issue roof (DESIGN.md "Roofline"):
// all cost arithmetic runs two pixels per 32-bit op in packed 16-bit lanes (u8 sums stay below
// 5355, so plain 32-bit adds never carry across lanes), |a-b| on pixels runs four per op
// (VABSDIFF4), the min over the nine costs and its tie-break are one VIMNMX3.U16x2 chain over
// (cost<<4 | rank) keys with the aa threshold folded in as a tenth key, and the interpolation
// operands are picked by a bitwise mux tree instead of branches.
//
// One thread owns 8 adjacent pool columns. Per pool row the block exchanges the vertical sums L
// through a double-buffered shared row (one __syncthreads per row); the running term
// M = B[r-1] + P[r] is the only state carried in registers (36 per thread).
//
// Reference semantics: /root/reference/src/SangNom2.cpp :60-65 (3-tap), :108-117 (costs),
// :138-152 (recursive blur, /16, wrap to u8), :208-249 (min, threshold, tie order, rounding mean).
#pragma once
#include "sangnom_cluster.cuh"
#include "sangnom_kernels.h"

#include <cstdint>

#ifndef SN_DYNAMIC_SMEM
#define SN_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace sn {
namespace u8k {

constexpr int kCols = 8;           // pool columns per thread
constexpr int kLPad = 8;           // u16 elements of padding on each side of a shared L row (16 B)

// rank of cost buffer i in the reference's tie order (4,5,3,6,2,7,1,8,0), in both 16-bit lanes
__device__ __forceinline__ constexpr uint32_t rank2(int i)
{
    constexpr int r[kNumCost] = { 8, 6, 4, 2, 0, 1, 3, 5, 7 };
    return (uint32_t)r[i] * 0x00010001u;
}

__device__ __forceinline__ uint32_t fsr(uint32_t lo, uint32_t hi, int bytes) { return __funnelshift_r(lo, hi, bytes * 8); }
__device__ __forceinline__ uint32_t lanes_lo(uint32_t w) { return __byte_perm(w, 0, 0x4140); }   // bytes 0,1 -> two u16 lanes
__device__ __forceinline__ uint32_t lanes_hi(uint32_t w) { return __byte_perm(w, 0, 0x4342); }   // bytes 2,3 -> two u16 lanes
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x6420); }   // low bytes of 4 lanes
__device__ __forceinline__ uint32_t absdiff2(uint32_t a, uint32_t b) { return __vmaxu2(a, b) - __vminu2(a, b); }
__device__ __forceinline__ uint32_t mean4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) & 0xFEFEFEFEu) >> 1); }   // (a+b+1)>>1 per byte
__device__ __forceinline__ uint32_t mux(uint32_t m, uint32_t a, uint32_t b) { return (a & ~m) | (b & m); }
// 0xFF in every byte whose bit `bit` is set: PRMT's sign-replicate selectors (8|byte index). __byte_perm
// masks the selector to 3 bits, so this goes through PTX directly.
__device__ __forceinline__ uint32_t byte_mask_from_bit(uint32_t ranks, int bit)
{
#ifdef SN_HOST_EMULATION
    return emul_prmt(ranks << (7 - bit), 0u, 0xBA98u);
#else
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(ranks << (7 - bit)), "r"(0u), "r"(0xBA98u));
    return d;
#endif
}

// 8 pixels at horizontal tap k (pixels x0+k .. x0+k+7) out of a window of bytes x0-4 .. x0+11.
struct Taps {
    uint32_t f1[3], f2[3], f3[3];   // window funnel-shifted right by 1, 2, 3 bytes
    uint32_t w1, w2;
    __device__ __forceinline__ void build(const uint32_t (&w)[4])
    {
#pragma unroll
        for (int i = 0; i < 3; ++i) { f1[i] = fsr(w[i], w[i + 1], 1); f2[i] = fsr(w[i], w[i + 1], 2); f3[i] = fsr(w[i], w[i + 1], 3); }
        w1 = w[1]; w2 = w[2];
    }
    // h = 0: pixels x0..x0+3 (+k), h = 1: pixels x0+4..x0+7 (+k)
    __device__ __forceinline__ uint32_t at(int k, int h) const
    {
        switch (k) {
            case -3: return f1[h];
            case -2: return f2[h];
            case -1: return f3[h];
            case 0: return h ? w2 : w1;
            case 1: return f1[h + 1];
            case 2: return f2[h + 1];
            default: return f3[h + 1];
        }
    }
};

// The two 3-tap values of one row for two pixels: f = T(m1,c,p1), b = T(p1,c,m1),
// T(a,c,d) = wrap8((4a + 5c - d) >> 3). The +2048 bias keeps every lane positive through the
// arithmetic shift and is a multiple of 256 after it, so it vanishes in the wrap.
__device__ __forceinline__ void tap3_pair(uint32_t m1, uint32_t c, uint32_t p1, uint32_t& f, uint32_t& b)
{
    const uint32_t u = c * 5u + 0x08000800u;
    f = (((m1 << 2) + u - p1) >> 3) & 0x00FF00FFu;
    b = (((p1 << 2) + u - m1) >> 3) & 0x00FF00FFu;
}

struct RowState {
    // 3-tap values of the pair (cur,next) kept for the interpolation one row later, as bytes
    uint32_t f1[2], f2[2], b1[2], b2[2];
};

// Window of bytes x0-4 .. x0+11 of a picture row with the reference's edge replication
// (loadPixel, SangNom2.cpp:25-34). vec: the row may be read with aligned 8/4-byte loads.
__device__ __forceinline__ void load_window(const uint8_t* __restrict__ row, int x0, int W, bool vec, uint32_t (&w)[4])
{
    if (x0 >= W) { w[0] = w[1] = w[2] = w[3] = 0; return; }
    if (vec) {
        const uint2 own = *reinterpret_cast<const uint2*>(row + x0);
        w[1] = own.x; w[2] = own.y;
        w[0] = x0 > 0 ? *reinterpret_cast<const uint32_t*>(row + x0 - 4) : 0u;
        w[3] = x0 + 8 < W ? *reinterpret_cast<const uint32_t*>(row + x0 + 8) : 0u;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int x = x0 - 4 + 4 * i + b;
                if (x >= 0 && x < W) v |= (uint32_t)row[x] << (8 * b);
            }
            w[i] = v;
        }
    }
    if (x0 == 0) w[0] = (w[1] & 0xFFu) * 0x01010101u;
    if (x0 + 11 > W - 1) {                       // right edge inside this window: replicate pixel W-1
        const int e = W - 1 - (x0 - 4);          // window byte index of the last picture pixel (>= 4)
        uint32_t ev = 0;
#pragma unroll
        for (int i = 1; i < 4; ++i) if ((e >> 2) == i) ev = (w[i] >> (8 * (e & 3))) & 0xFFu;
        ev *= 0x01010101u;
#pragma unroll
        for (int i = 1; i < 4; ++i) {
            const int first = 4 * i;             // window byte index of this word's byte 0
            if (e < first) w[i] = ev;
            else if (e < first + 3) { const uint32_t keep = 0xFFFFFFFFu >> (8 * (3 - (e - first))); w[i] = (w[i] & keep) | (ev & ~keep); }
        }
    }
}

// Eight stale cost bytes of buffer i at pool row r, columns x0..x0+7 (0 outside the handed-over regions).
__device__ __forceinline__ uint2 state_load8(const CostState& s, int i, int r, int x0, int S)
{
    if (s.b != nullptr && r >= s.b_r0 && r <= s.b_r1) {
        const int nb = s.b_r1 - s.b_r0 + 1;
        return __ldg(reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(s.b) + ((size_t)i * nb + (r - s.b_r0)) * S + x0));
    }
    if (s.a != nullptr && x0 >= s.a_x0 && r >= 1 && r <= s.a_rows) {
        const int wa = S - s.a_x0;
        return __ldg(reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(s.a) + ((size_t)i * (s.a_rows + 1) + r) * wa + (x0 - s.a_x0)));
    }
    return make_uint2(0u, 0u);
}

__device__ __forceinline__ void state_store8(const CostState& s, int i, int r, int x0, int S, uint2 v)
{
    if (s.b != nullptr && r >= s.b_r0 && r <= s.b_r1) {
        const int nb = s.b_r1 - s.b_r0 + 1;
        *reinterpret_cast<uint2*>(static_cast<uint8_t*>(s.b) + ((size_t)i * nb + (r - s.b_r0)) * S + x0) = v;
    } else if (s.a != nullptr && x0 >= s.a_x0 && r >= 1 && r <= s.a_rows) {
        const int wa = S - s.a_x0;
        *reinterpret_cast<uint2*>(static_cast<uint8_t*>(s.a) + ((size_t)i * (s.a_rows + 1) + r) * wa + (x0 - s.a_x0)) = v;
    }
}

// Raw costs P of one row pair for this thread's 8 pixels, as nine sets of four packed-lane words,
// plus the 3-tap values the interpolation of this pair will need.
__device__ __forceinline__ void pair_costs(const Taps& c, const Taps& n, uint32_t (&P)[kNumCost][4], RowState& keep)
{
    // seven pixel-pair costs, four pixels per VABSDIFF4, then widened to 16-bit lanes
    const int tap_of[kNumCost] = { -3, -2, -1, 0, 0, 0, 1, 2, 3 };
#pragma unroll
    for (int i = 0; i < kNumCost; ++i) {
        if (i == 3 || i == 5) continue;
        const int k = tap_of[i];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t d = __vabsdiffu4(c.at(k, h), n.at(-k, h));
            P[i][2 * h] = lanes_lo(d);
            P[i][2 * h + 1] = lanes_hi(d);
        }
    }
    // the two 3-tap costs, two pixels per op
    uint32_t f1[4], b1[4], f2[4], b2[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t cm = c.at(-1, h), cc = c.at(0, h), cp = c.at(1, h);
        const uint32_t nm = n.at(-1, h), nc = n.at(0, h), np = n.at(1, h);
        tap3_pair(lanes_lo(cm), lanes_lo(cc), lanes_lo(cp), f1[2 * h], b1[2 * h]);
        tap3_pair(lanes_hi(cm), lanes_hi(cc), lanes_hi(cp), f1[2 * h + 1], b1[2 * h + 1]);
        tap3_pair(lanes_lo(nm), lanes_lo(nc), lanes_lo(np), b2[2 * h], f2[2 * h]);      // next row: roles swap
        tap3_pair(lanes_hi(nm), lanes_hi(nc), lanes_hi(np), b2[2 * h + 1], f2[2 * h + 1]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { P[3][q] = absdiff2(f1[q], f2[q]); P[5][q] = absdiff2(b1[q], b2[q]); }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        keep.f1[h] = pack4(f1[2 * h], f1[2 * h + 1]); keep.f2[h] = pack4(f2[2 * h], f2[2 * h + 1]);
        keep.b1[h] = pack4(b1[2 * h], b1[2 * h + 1]); keep.b2[h] = pack4(b2[2 * h], b2[2 * h + 1]);
    }
}

// Interpolated pixels (8 bytes) of the row between `c` and `n` from the four min-key words.
__device__ __forceinline__ uint2 interpolate8(const Taps& c, const Taps& n, const RowState& sg, const uint32_t (&kmin)[4])
{
    uint2 out;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t ranks = pack4(kmin[2 * h], kmin[2 * h + 1]);     // low nibble of each byte = winning rank
        const uint32_t m0 = byte_mask_from_bit(ranks, 0), m1 = byte_mask_from_bit(ranks, 1);
        const uint32_t m2 = byte_mask_from_bit(ranks, 2), m3 = byte_mask_from_bit(ranks, 3);
        // operands by rank: 0 (c0,n0) 1 (b1,b2) 2 (f1,f2) 3 (c+1,n-1) 4 (c-1,n+1) 5 (c+2,n-2) 6 (c-2,n+2) 7 (c+3,n-3) 8 (c-3,n+3)
        const uint32_t a01 = mux(m0, c.at(0, h), sg.b1[h]), a23 = mux(m0, sg.f1[h], c.at(1, h));
        const uint32_t a45 = mux(m0, c.at(-1, h), c.at(2, h)), a67 = mux(m0, c.at(-2, h), c.at(3, h));
        const uint32_t a = mux(m3, mux(m2, mux(m1, a01, a23), mux(m1, a45, a67)), c.at(-3, h));
        const uint32_t b01 = mux(m0, n.at(0, h), sg.b2[h]), b23 = mux(m0, sg.f2[h], n.at(-1, h));
        const uint32_t b45 = mux(m0, n.at(1, h), n.at(-2, h)), b67 = mux(m0, n.at(2, h), n.at(-3, h));
        const uint32_t b = mux(m3, mux(m2, mux(m1, b01, b23), mux(m1, b45, b67)), n.at(3, h));
        (h ? out.y : out.x) = mean4(a, b);
    }
    return out;
}

template <int kMaxThreads, int kMinBlocks, bool kClustered>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks)
sangnom_u8_row_sweep(const PlaneTask* __restrict__ tasks, LaunchGeometry g, int seg_cols)
{
    SN_DYNAMIC_SMEM(smem_raw);
    // a plane wider than one block is split into column segments over the blocks of a cluster
    const unsigned G = kClustered ? cl::size() : 1u;      // kClustered = false: one block per plane, no cluster code at all
    const unsigned crank = kClustered ? cl::rank() : 0u;
    const PlaneTask t = tasks[blockIdx.x / G];
    const int S = g.S;
    const int LS = seg_cols + 2 * kLPad;                            // u16 elements per shared L row of this segment
    uint16_t* const Lbase = reinterpret_cast<uint16_t*>(smem_raw);  // [2][9][LS]

    const int W = t.width, n = t.kept_rows, R = t.sweep_rows;
    const int lx = threadIdx.x * kCols;                             // column inside the segment
    const int x0 = (int)crank * seg_cols + lx;                      // pool column
    const bool plane_first = x0 == 0, plane_last = x0 + kCols == S;
    const bool seg_first = lx == 0, seg_last = lx + kCols == seg_cols;
    uint8_t* const plane = static_cast<uint8_t*>(t.plane);
    const uint8_t* const src = static_cast<const uint8_t*>(t.src);
    const long long pitch = t.pitch, src_pitch = t.src_pitch;
    const long long wpad = ((long long)W + 15) & ~15LL;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)src_pitch) & 15) == 0 && src_pitch >= wpad;      // aligned vector loads of kept rows
    const bool vec_out = ((reinterpret_cast<uintptr_t>(plane) | (uintptr_t)pitch) & 15) == 0 && pitch >= wpad;        // aligned vector stores
    // which of my 8 columns carry pixels: all, none, or a prefix (the one thread that straddles W)
    const int npix = min(max(W - x0, 0), kCols);
    const uint32_t pixmask_lo = npix >= 4 ? 0xFFFFFFFFu : (npix <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - npix))));
    const uint32_t pixmask_hi = npix >= 8 ? 0xFFFFFFFFu : (npix <= 4 ? 0u : (0xFFFFFFFFu >> (8 * (8 - npix))));

    auto kept_row = [&](int j) -> const uint8_t* { return src + (long long)j * src_pitch; };
    auto store8 = [&](uint8_t* row, uint2 v) {
        if (npix == kCols && vec_out) { *reinterpret_cast<uint2*>(row + x0) = v; return; }
#pragma unroll
        for (int b = 0; b < 8; ++b) if (b < npix) row[x0 + b] = (uint8_t)(((b < 4 ? v.x : v.y) >> (8 * (b & 3))) & 0xFFu);
    };

    // ---- border row without a neighbour pair (reference GetFrame :380-391) ----
    if (npix > 0) {
        uint8_t* to = t.offset == 0 ? plane + (long long)(t.height - 1) * pitch : plane;
        uint32_t w[4];
        load_window(kept_row(t.offset == 0 ? n - 1 : 0), x0, W, vec, w);
        store8(to, make_uint2(w[1], w[2]));
        if (t.copy_kept) {                       // last kept row; rows 0..n-2 are written as the sweep passes them
            if (t.offset != 0) load_window(kept_row(n - 1), x0, W, vec, w);
            store8(plane + (long long)(t.offset + 2 * (n - 1)) * pitch, make_uint2(w[1], w[2]));
        }
    }

    // ---- running term M = B[r-1] + P[r]; B[0] = 0 so M starts as P[1] ----
    uint32_t M[kNumCost][4];
    uint32_t wa[4], wb[4], wc[4], wpre[4];          // windows of K[r-1], K[r], K[r+1], K[r+2]
    RowState sg_prev{}, sg_next{};                   // 3-tap bytes of pairs (r-1,r) and (r,r+1)

    // Raw cost row `row` (pool row index) into P: from pixels where this thread has them and the
    // pair exists, otherwise from the cost state the previous pass of the frame left.
    auto cost_row = [&](int row, const uint32_t (&cw)[4], const uint32_t (&nw)[4], uint32_t (&P)[kNumCost][4], RowState& keep) {
        const bool pair = row <= n - 1;              // pool row j+1 holds the costs of pair (K[j], K[j+1])
        if (pair && npix == kCols) {
            Taps c, nx;
            c.build(cw); nx.build(nw);
            pair_costs(c, nx, P, keep);
            return;
        }
        uint2 st[kNumCost];
#pragma unroll
        for (int i = 0; i < kNumCost; ++i) st[i] = state_load8(t.in, i, row, x0, S);
        if (pair && npix > 0) {                      // the straddling thread: pixels left, stale right
            Taps c, nx;
            c.build(cw); nx.build(nw);
            pair_costs(c, nx, P, keep);
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                const uint32_t lo = (pack4(P[i][0], P[i][1]) & pixmask_lo) | (st[i].x & ~pixmask_lo);
                const uint32_t hi = (pack4(P[i][2], P[i][3]) & pixmask_hi) | (st[i].y & ~pixmask_hi);
                P[i][0] = lanes_lo(lo); P[i][1] = lanes_hi(lo); P[i][2] = lanes_lo(hi); P[i][3] = lanes_hi(hi);
            }
        } else {
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                P[i][0] = lanes_lo(st[i].x); P[i][1] = lanes_hi(st[i].x); P[i][2] = lanes_lo(st[i].y); P[i][3] = lanes_hi(st[i].y);
            }
        }
    };

    load_window(kept_row(0), x0, W, vec, wa);
    if (n >= 2) load_window(kept_row(1), x0, W, vec, wb); else { wb[0] = wb[1] = wb[2] = wb[3] = 0; }
    if (n >= 3) load_window(kept_row(2), x0, W, vec, wc); else { wc[0] = wc[1] = wc[2] = wc[3] = 0; }
    cost_row(1, wa, wb, M, sg_next);

    const uint32_t tkey = (uint32_t)min(t.thr_i + 1, 4095) * 0x00100010u;    // (thr+1) << 4 in both lanes
    const bool exporting = t.out.a != nullptr || t.out.b != nullptr;

    for (int r = 1; r <= R; ++r) {
        // wa = K[r-1], wb = K[r], wc = K[r+1]; start the loads of K[r+2]
        if (r + 2 <= n - 1) load_window(kept_row(r + 2), x0, W, vec, wpre);

        // ---- P[r+1], L = M + P[r+1] -> shared row; M keeps P[r+1] until B[r] is known ----
        sg_prev = sg_next;
        uint16_t* const Lrow = Lbase + (size_t)(r & 1) * kNumCost * LS + kLPad;
        {
            uint32_t P[kNumCost][4];
            cost_row(r + 1, wb, wc, P, sg_next);
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                uint4 L;
                L.x = M[i][0] + P[i][0]; L.y = M[i][1] + P[i][1]; L.z = M[i][2] + P[i][2]; L.w = M[i][3] + P[i][3];
                uint16_t* row = Lrow + i * LS;
                *reinterpret_cast<uint4*>(row + lx) = L;
                if (seg_first) {
                    if (plane_first) { const uint32_t e = (L.x & 0xFFFFu) * 0x00010001u; *reinterpret_cast<uint2*>(row - 4) = make_uint2(e, e); }   // clamp at column 0
                    else {                                                       // my first columns are the left neighbour's right halo
                        cl::store_remote(reinterpret_cast<uint32_t*>(row + seg_cols), crank - 1, L.x);
                        cl::store_remote(reinterpret_cast<uint32_t*>(row + seg_cols + 2), crank - 1, L.y);
                    }
                }
                if (seg_last) {
                    if (plane_last) { const uint32_t e = (L.w >> 16) * 0x00010001u; *reinterpret_cast<uint2*>(row + seg_cols) = make_uint2(e, e); }   // clamp at column S-1
                    else {                                                       // my last columns are the right neighbour's left halo
                        cl::store_remote(reinterpret_cast<uint32_t*>(row - 4), crank + 1, L.z);
                        cl::store_remote(reinterpret_cast<uint32_t*>(row - 2), crank + 1, L.w);
                    }
                }
                M[i][0] = P[i][0]; M[i][1] = P[i][1]; M[i][2] = P[i][2]; M[i][3] = P[i][3];
            }
        }
        if constexpr (kClustered) cl::sync_all(); else __syncthreads();

        // ---- B[r] = wrap8(H7(L) >> 4) per buffer; keys for the min; M += B ----
        uint32_t kmin[4] = { tkey, tkey, tkey, tkey };
#pragma unroll
        for (int i = 0; i < kNumCost; ++i) {
            const uint16_t* row = Lrow + i * LS + lx;
            const uint2 lh = *reinterpret_cast<const uint2*>(row - 4);      // (l-4,l-3) (l-2,l-1)
            const uint4 own = *reinterpret_cast<const uint4*>(row);         // (l0,l1) .. (l6,l7)
            const uint2 rh = *reinterpret_cast<const uint2*>(row + 8);      // (l8,l9) (l10,l11)
            const uint32_t Wm2 = lh.x, Wm1 = lh.y, W0 = own.x, W1 = own.y, W2 = own.z, W3 = own.w, W4 = rh.x, W5 = rh.y;
            // Z_k = W[k-1]+W[k]+W[k+1] (even / odd triples), X_k = Z_k + W[k-2];
            // H7_k = Z_k + (X_k.hi, X_{k+1}.lo)  -> lanes (sum l[2k-3..2k+3], sum l[2k-2..2k+4])
            const uint32_t Z0 = Wm1 + W0 + W1, Z1 = W0 + W1 + W2, Z2 = W1 + W2 + W3, Z3 = W2 + W3 + W4, Z4 = W3 + W4 + W5;
            const uint32_t X0 = Z0 + Wm2, X1 = Z1 + Wm1, X2 = Z2 + W0, X3 = Z3 + W1, X4 = Z4 + W2;
            uint32_t H[4];
            H[0] = Z0 + __funnelshift_r(X0, X1, 16);
            H[1] = Z1 + __funnelshift_r(X1, X2, 16);
            H[2] = Z2 + __funnelshift_r(X2, X3, 16);
            H[3] = Z3 + __funnelshift_r(X3, X4, 16);
            uint32_t Bq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                Bq[q] = (H[q] >> 4) & 0x00FF00FFu;
                M[i][q] += Bq[q];
                const uint32_t key = (H[q] & 0x0FF00FF0u) | rank2(i);
                kmin[q] = __vminu2(kmin[q], key);
            }
            // hand the blurred row to the next pass of this frame
            if (exporting) state_store8(t.out, i, r, x0, S, make_uint2(pack4(Bq[0], Bq[1]), pack4(Bq[2], Bq[3])));
        }

        // ---- interpolate the picture row between K[r-1] and K[r] ----
        if (r <= n - 1 && npix > 0) {
            Taps c, nx;
            c.build(wa); nx.build(wb);
            const uint2 px = interpolate8(c, nx, sg_prev, kmin);
            store8(plane + (long long)(t.offset + 2 * (r - 1) + 1) * pitch, px);
            if (t.copy_kept) store8(plane + (long long)(t.offset + 2 * (r - 1)) * pitch, make_uint2(wa[1], wa[2]));
        }

#pragma unroll
        for (int q = 0; q < 4; ++q) { wa[q] = wb[q]; wb[q] = wc[q]; wc[q] = wpre[q]; }
    }
}

inline size_t smem_bytes(int seg_cols) { return (size_t)2 * kNumCost * (seg_cols + 2 * kLPad) * sizeof(uint16_t); }

}  // namespace u8k
}  // namespace sn
