// 8-bit flavour of the fused row sweep, written for the instruction-issue roof.
//
// One thread owns 8 adjacent pool columns. Per pool row r:
//   phase A  kept row K[r+1] arrives in a shared-memory ring (bulk async copy, sangnom_stage.cuh); the thread
//            takes its 16-byte window, builds the +-3 byte shifts, the 3-tap values of the row (once per
//            row, reused for two pairs), the nine raw costs P[r+1] four pixels per VABSDIFF4, widens them
//            to 16-bit lanes and publishes L = B[r-1] + P[r] + P[r+1] to a double-buffered shared row;
//   barrier  one block barrier per row; when the plane is split over the blocks of a cluster, the two edge threads
//            of a block also exchange their halo with the neighbour blocks (sangnom_cluster.cuh) - no cluster barrier;
//   phase B  per cost: 7-tap sum of L two columns per op, key = (sum & 0x0FF0) | rank in both lanes
//            ((B << 4) | tie-break rank; B = wrap8(sum >> 4)), running term M = P[r+1] + (key >> 4) in ONE
//            LEA.HI - the rank nibble of the upper lane that this shifts into the lower lane is a per-cost
//            constant and is taken out again by the immediate of the next row's IADD3 - and a
//            VIMNMX3.U16x2 chain over the keys with the aa threshold as a tenth key; then the interpolated
//            picture row between K[r-1] and K[r] through a bitwise mux tree.
// State carried in registers: M (36), the windows of three kept rows (12), their 3-tap bytes (12).
//
// Reference semantics: /root/reference/src/SangNom2.cpp :25-34 (edge replication), :60-65 (3-tap),
// :108-117 (costs), :138-152 (recursive cost sum, /16, wrap to u8), :208-249 (min, threshold, tie order,
// rounding mean), GetFrame :361-391 (kept field, border row).
#pragma once
#include "sangnom_cluster.cuh"
#include "sangnom_kernels.h"
#include "sangnom_stage.cuh"

#include <cstdint>
#include <type_traits>

#ifndef SN_DYNAMIC_SMEM
#define SN_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace sn {
namespace u8k {

constexpr int kCols = 8;           // pool columns per thread
constexpr int kLEntry = 2 * kNumCost + 1;   // uint2 words per thread entry of the shared L rows: 9 costs x 2 halves + 1 pad (152 B: conflict-free)
constexpr int kT3Ring = 4;         // 3-tap byte ring slots (thread-private)
constexpr int kRing = 8;           // kept-row ring slots: rows r-1 .. r+1 in use, up to r+kAhead in flight
constexpr int kAhead = 4;          // rows staged ahead of the row being consumed
constexpr int kRingPad = 16;       // bytes of halo on each side of a staged row segment

// rank of cost buffer i in the reference's tie order (4,5,3,6,2,7,1,8,0)
__device__ __forceinline__ constexpr uint32_t rank1(int i)
{
    constexpr int r[kNumCost] = { 8, 6, 4, 2, 0, 1, 3, 5, 7 };
    return (uint32_t)r[i];
}
__device__ __forceinline__ constexpr uint32_t rank2(int i) { return rank1(i) * 0x00010001u; }   // in both 16-bit lanes
// what (key >> 4) leaks from the upper lane's rank nibble into the lower lane
__device__ __forceinline__ constexpr uint32_t leak(int i) { return rank1(i) << 12; }

// a value the compiler cannot see through: what is computed from it stays where it is written (no hoisting of a
// cold path's loop-invariant arithmetic into the hot loop)
__device__ __forceinline__ int opaque(int v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ uint32_t fsr(uint32_t lo, uint32_t hi, int bytes) { return __funnelshift_r(lo, hi, bytes * 8); }
__device__ __forceinline__ uint32_t lanes_lo(uint32_t w) { return __byte_perm(w, 0, 0x4140); }   // bytes 0,1 -> two u16 lanes
__device__ __forceinline__ uint32_t lanes_hi(uint32_t w) { return __byte_perm(w, 0, 0x4342); }   // bytes 2,3 -> two u16 lanes
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x6420); }   // low bytes of 4 lanes
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

// The two 3-tap values of every pixel of one kept row, as bytes: f = T(x-1, x, x+1), b = T(x+1, x, x-1),
// T(a,c,d) = wrap8((4a + 5c - d) >> 3). Computed once per row: as the upper row of a pair it supplies
// (f1, b1) = (f, b), as the lower row (f2, b2) = (b, f) (reference :103-106).
struct Tap3 { uint32_t f[2], b[2]; };

// kSat: the arithmetic of the reference's SSE2 path (SangNom2_SSE2.cpp:449-481) - 16-bit unsigned lanes, logical
// shift, pack with saturation: a negative sum becomes 255, a quotient above 255 clamps.
template <bool kSat>
__device__ __forceinline__ void tap3_row(const Taps& t, Tap3& o)
{
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t m = t.at(-1, h), c = t.at(0, h), p = t.at(1, h);
        uint32_t fl[2], bl[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t m2 = q ? lanes_hi(m) : lanes_lo(m), c2 = q ? lanes_hi(c) : lanes_lo(c), p2 = q ? lanes_hi(p) : lanes_lo(p);
            // the +2048 bias keeps every lane positive through the shift and is a multiple of 256 after it,
            // so it vanishes in the wrap; the bits the 32-bit shift drags across lanes stay above bit 12
            const uint32_t u = c2 * 5u + 0x08000800u;
            fl[q] = ((m2 << 2) + u - p2) >> 3;
            bl[q] = ((p2 << 2) + u - m2) >> 3;
            if constexpr (kSat) {
                // lane value = (sum >> 3) + 256 in [224, 542]; w = value - 224: w < 32 <=> sum < 0 -> 255,
                // else min(w - 32, 255)
                auto sat = [](uint32_t v) {
                    const uint32_t w = (v & 0x03FF03FFu) - 0x00E000E0u;
                    const uint32_t z = __vminu2(__vmaxu2(w, 0x00200020u) - 0x00200020u, 0x00FF00FFu);
                    const uint32_t neg = (((0x00200020u - __vminu2(w, 0x00200020u)) + 0x00FF00FFu) >> 8) & 0x00010001u;
                    return z | (neg * 0xFFu);
                };
                fl[q] = sat(fl[q]);
                bl[q] = sat(bl[q]);
            }
        }
        o.f[h] = pack4(fl[0], fl[1]);
        o.b[h] = pack4(bl[0], bl[1]);
    }
}

// Nine raw costs of the pair (upper row c, lower row n) for 8 pixels, as bytes (two words per cost).
__device__ __forceinline__ void pair_costs(const Taps& c, const Tap3& c3, const Taps& n, const Tap3& n3, uint32_t (&P)[kNumCost][2])
{
    const int tap_of[kNumCost] = { -3, -2, -1, 0, 0, 0, 1, 2, 3 };
#pragma unroll
    for (int i = 0; i < kNumCost; ++i) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (i == 3) P[i][h] = __vabsdiffu4(c3.f[h], n3.b[h]);          // |f1 - f2|
            else if (i == 5) P[i][h] = __vabsdiffu4(c3.b[h], n3.f[h]);     // |b1 - b2|
            else P[i][h] = __vabsdiffu4(c.at(tap_of[i], h), n.at(-tap_of[i], h));
        }
    }
}

// Interpolated pixels (8 bytes) of the row between `c` and `n` from the four min-key words.
__device__ __forceinline__ uint2 interpolate8(const Taps& c, const Tap3& c3, const Taps& n, const Tap3& n3, const uint32_t (&kmin)[4])
{
    uint2 out;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t ranks = pack4(kmin[2 * h], kmin[2 * h + 1]);     // low nibble of each byte = winning rank
        const uint32_t m0 = byte_mask_from_bit(ranks, 0), m1 = byte_mask_from_bit(ranks, 1);
        const uint32_t m2 = byte_mask_from_bit(ranks, 2), m3 = byte_mask_from_bit(ranks, 3);
        // operands by rank: 0 (c0,n0) 1 (b1,b2) 2 (f1,f2) 3 (c+1,n-1) 4 (c-1,n+1) 5 (c+2,n-2) 6 (c-2,n+2) 7 (c+3,n-3) 8 (c-3,n+3)
        const uint32_t a01 = mux(m0, c.at(0, h), c3.b[h]), a23 = mux(m0, c3.f[h], c.at(1, h));
        const uint32_t a45 = mux(m0, c.at(-1, h), c.at(2, h)), a67 = mux(m0, c.at(-2, h), c.at(3, h));
        const uint32_t a = mux(m3, mux(m2, mux(m1, a01, a23), mux(m1, a45, a67)), c.at(-3, h));
        const uint32_t b01 = mux(m0, n.at(0, h), n3.f[h]), b23 = mux(m0, n3.b[h], n.at(-1, h));
        const uint32_t b45 = mux(m0, n.at(1, h), n.at(-2, h)), b67 = mux(m0, n.at(2, h), n.at(-3, h));
        const uint32_t b = mux(m3, mux(m2, mux(m1, b01, b23), mux(m1, b45, b67)), n.at(3, h));
        (h ? out.y : out.x) = mean4(a, b);
    }
    return out;
}

// Where the cost state of one pool row lives for this thread's 8 columns: bytes of buffer i at p + i * stride.
// p == nullptr: outside the handed-over regions (reads as the zero-filled pool, nothing to write).
struct StateRow { uint8_t* p; size_t stride; };

__device__ __forceinline__ StateRow state_row(const CostState& s, int r, int x0, int S)
{
    if (s.b != nullptr && r >= s.b_r0 && r <= s.b_r1) {
        const int nb = s.b_r1 - s.b_r0 + 1;
        return StateRow{ static_cast<uint8_t*>(s.b) + (size_t)(r - s.b_r0) * S + x0, (size_t)nb * S };
    }
    if (s.a != nullptr && x0 >= s.a_x0 && r >= 1 && r <= s.a_rows) {
        const int wa = S - s.a_x0;
        return StateRow{ static_cast<uint8_t*>(s.a) + (size_t)r * wa + (x0 - s.a_x0), (size_t)(s.a_rows + 1) * wa };
    }
    return StateRow{ nullptr, 0 };
}

// 8 bytes at row positions p0 .. p0+7 with zero outside [0, W); `fast`: 8-byte aligned loads are legal
__device__ __forceinline__ uint2 load8_guarded(const uint8_t* __restrict__ row, int p0, int W, bool fast)
{
    if (fast && p0 >= 0 && p0 + 8 <= W) return *reinterpret_cast<const uint2*>(row + p0);
    uint2 v = make_uint2(0u, 0u);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int p = p0 + b;
        if (p >= 0 && p < W) { const uint32_t px = (uint32_t)row[p] << (8 * (b & 3)); if (b < 4) v.x |= px; else v.y |= px; }
    }
    return v;
}

// Shared memory of one block (seg_cols pool columns, T = seg_cols / 8 threads):
//   L     [2 parities][T + 2 entries][9 costs][2 halves] uint2 (+1 pad word per entry): the vertical sums of a
//         thread's 8 columns as 16-bit lanes, half 0 = columns 0..3, half 1 = columns 4..7; entry 0 and T+1 hold the
//         neighbour segment's edge or the clamp. Entries are 152 bytes apart, so consecutive lanes hit distinct
//         banks with 8-byte accesses and every per-cost offset is an immediate.
//   ring  [kRing][ring_stride]  staged kept rows, row position p at offset p - seg_x0 + kRingPad
//   t3    [kT3Ring][T] uint4    3-tap bytes of the kept rows (thread-private slots)
//   mbar  [kRing]               one mbarrier per ring slot
//   halo  [2 sides][2 parities] barriers the neighbour blocks' halo stores complete on (cluster launches)
//   task                        this block's PlaneTask
inline __host__ __device__ int ring_stride(int seg_cols) { return (seg_cols + 2 * kRingPad + 15) & ~15; }
inline __host__ __device__ size_t l_bytes(int seg_cols) { return (size_t)2 * (seg_cols / kCols + 2) * kLEntry * sizeof(uint2); }
inline size_t smem_bytes(int seg_cols)
{
    return ((l_bytes(seg_cols) + 15) & ~(size_t)15) + (size_t)kRing * ring_stride(seg_cols) + (size_t)kT3Ring * (seg_cols / kCols) * sizeof(uint4) +
           kRing * sizeof(stage::Mbar) + 4 * 16 /* halo barriers */ + ((sizeof(PlaneTask) + 15) & ~(size_t)15);
}

// kSpare: the launch brings spare threads for planes narrower than the pool (see the thread -> column map below); planes
// as wide as the pool are launched without (kSpare = false: threadIdx.x is the working thread index, nothing else).
template <int kMaxThreads, int kMinBlocks, bool kClustered, bool kSat = false, bool kSpare = false>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks)
sangnom_u8_row_sweep(const PlaneTask* __restrict__ tasks, LaunchGeometry g, int seg_cols)
{
    SN_DYNAMIC_SMEM(smem_raw);
    // a plane wider than one block is split into column segments over the blocks of a cluster
    const unsigned G = kClustered ? cl::size() : 1u;      // kClustered = false: one block per plane, no cluster code at all
    const unsigned crank = kClustered ? cl::rank() : 0u;
    const int S = g.S;
    const uint32_t keymask = g.key_mask;                            // 0x0FF00FF0, kept in a register so (sum & mask) | rank is one LOP3
    const int T = seg_cols / kCols;                                 // working threads of the block
    uint2* const Lbase = reinterpret_cast<uint2*>(smem_raw);
    const int rstride = ring_stride(seg_cols);
    uint8_t* const ring = smem_raw + ((l_bytes(seg_cols) + 15) & ~(size_t)15);
    uint4* const t3ring = reinterpret_cast<uint4*>(ring + (size_t)kRing * rstride);
    stage::Mbar* const mbar = reinterpret_cast<stage::Mbar*>(t3ring + (size_t)kT3Ring * T);
    // the task lives in shared memory: its rarely used fields (cost-state regions, pitches) are re-read where needed
    // instead of occupying registers for the whole sweep
    // halo barriers, 16 bytes apart: [side 0 = left, 1 = right][row parity]
    unsigned char* const halo_raw = reinterpret_cast<unsigned char*>(mbar + kRing);
    auto halo_bar = [&](int side, int par) -> cl::HaloBar* { return reinterpret_cast<cl::HaloBar*>(halo_raw + (side * 2 + par) * 16); };
    unsigned char* const task_raw = halo_raw + 4 * 16;
    {
        uint32_t* const dst = reinterpret_cast<uint32_t*>(task_raw);
        const uint32_t* const from = reinterpret_cast<const uint32_t*>(tasks + blockIdx.x / G);
        for (unsigned k = threadIdx.x; k < sizeof(PlaneTask) / 4; k += blockDim.x) dst[k] = from[k];
        __syncthreads();
    }
    const PlaneTask& t = *reinterpret_cast<const PlaneTask*>(task_raw);

    const int W = t.width, n = t.kept_rows, R = t.sweep_rows;
    const int seg_x0 = (int)crank * seg_cols;
    // Thread -> column map. A plane narrower than the pool (subsampled chroma in the luma-wide pool) ends its pixel
    // threads inside a warp; that warp would carry pixel lanes AND state-only lanes (hand-over loads and stores), run
    // both paths one after the other and be the slowest warp of the block, the one every row barrier waits for. The
    // launch therefore brings spare threads, and the first `shift` threads of the block are left idle so that the
    // last pixel thread ends a warp and the state-only threads start the next one. Spare threads leave at once.
    const int hw = (int)threadIdx.x;
    const int Tpx = (min(max(W - seg_x0, 0), seg_cols) + kCols - 1) / kCols;       // threads of this segment that carry pixels
    int shift = (kSpare && !kClustered && Tpx > 0 && Tpx < T && (Tpx & 31) != 0) ? 32 - (Tpx & 31) : 0;
    if (T + shift > (int)blockDim.x) shift = 0;
    const int tid = kSpare ? hw - shift : hw;                       // working thread index (block-uniform offset)
    if (kSpare && (tid < 0 || tid >= T)) {                          // spare thread: nothing to do, not even the barriers
#ifdef SN_HOST_EMULATION
        emul::bar->arrive_and_drop();
#endif
        return;
    }
    // (spare lanes have exited: the warp votes below may still name them in their mask)
    const int wfirst = max((hw & ~31) - shift, 0);                  // first working thread of my warp
#ifdef SN_HOST_EMULATION
    const int wlast = min((hw & ~31) + 31 - shift, T - 1);          // the emulation has no warps: working thread range of my "warp"
#endif
    const int lx = tid * kCols;                                     // column inside the segment
    const int x0 = seg_x0 + lx;                                     // pool column
    const bool plane_first = x0 == 0, plane_last = x0 + kCols == S;
    const bool seg_first = tid == 0, seg_last = tid == T - 1;
    const uint8_t* const src = static_cast<const uint8_t*>(t.src);
    const long long src_pitch = t.src_pitch;
    const int wpad = (W + 15) & ~15;
    // which of my 8 columns carry pixels: all, none, or a prefix (the one thread that straddles W)
    const int npix = min(max(W - x0, 0), kCols);
    // Picture edges (reference loadPixel :25-34: pixel 0 and pixel W-1 replicated outwards) are made in the staged rows
    // themselves, once per kept row, by one thread each - not in every window that touches them (three per pool row):
    // the block that holds pool column 0 fills positions -4..-1, the block(s) with a pixel thread whose window reaches
    // past W-1 fill W..W+10. Done two rows ahead of a row's first use, so the row barrier in between publishes it.
    const bool patch_left = seg_x0 == 0 && tid == 0 && W > 0;
    const bool patch_right = seg_x0 < W && W <= seg_x0 + seg_cols + 3 && tid == min(T - 1, (W - 1 - seg_x0) / kCols);
    const bool vec_out = ((reinterpret_cast<uintptr_t>(t.plane) | (uintptr_t)t.pitch) & 7) == 0 && npix == kCols;   // aligned 8-byte stores


    // ---- staging of kept rows: positions [lo, hi) of every kept row go to ring offset (position - seg_x0 + 16) ----
    const int lo = max(seg_x0 - kRingPad, 0), hi = min(seg_x0 + seg_cols + kRingPad, wpad);
    const bool seg_has_pixels = seg_x0 < W;
    // bulk copies need 16-byte aligned rows and may read up to wpad bytes of a row
    const bool bulk = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)src_pitch | (uintptr_t)seg_cols) & 15) == 0 && src_pitch >= wpad;
    const bool fast8 = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)src_pitch) & 7) == 0;
    auto kept_row = [&](int j) -> const uint8_t* { return src + (long long)j * src_pitch; };
    auto slot_of = [&](int j) -> uint8_t* { return ring + (size_t)(j & (kRing - 1)) * rstride; };
    // bulk mode: thread 0 starts the copy of kept row j
    auto issue_row = [&](int j) {
        stage::bulk_load(slot_of(j) + (lo - seg_x0 + kRingPad), kept_row(j) + lo, (unsigned)(hi - lo), &mbar[j & (kRing - 1)]);
    };
    // cooperative mode (unaligned sources): my own 8 bytes of a row, edge threads also the neighbour segments' halo
    auto coop_load = [&](int j, uint2 (&v)[3]) {
        const uint8_t* row = kept_row(j);
        const int xo = opaque(x0);      // keeps this rare path's address arithmetic inside its branch, out of every row's head
        v[0] = load8_guarded(row, xo, W, fast8);
        if (kClustered && seg_first) v[1] = load8_guarded(row, xo - 8, W, fast8);
        if (kClustered && seg_last) v[2] = load8_guarded(row, xo + 8, W, fast8);
    };
    auto coop_store = [&](int j, const uint2 (&v)[3]) {
        uint8_t* s = slot_of(j) + kRingPad + lx;
        *reinterpret_cast<uint2*>(s) = v[0];
        if (kClustered && seg_first) *reinterpret_cast<uint2*>(s - 8) = v[1];
        if (kClustered && seg_last) *reinterpret_cast<uint2*>(s + 8) = v[2];
    };
    // wait until kept row j has landed in the ring (first use of a row only)
    auto await_row = [&](int j) { if (bulk) stage::mbar_wait(&mbar[j & (kRing - 1)], (unsigned)(j / kRing) & 1u); };
    // my window of kept row j (bytes x0-4 .. x0+11) out of the ring, picture edges replicated
    auto window = [&](int j, uint32_t (&w)[4]) {
        const uint8_t* s = slot_of(j) + kRingPad + lx;
        w[0] = *reinterpret_cast<const uint32_t*>(s - 4);
        const uint2 own = *reinterpret_cast<const uint2*>(s);
        w[1] = own.x; w[2] = own.y;
        w[3] = *reinterpret_cast<const uint32_t*>(s + 8);
    };
    // the edge replication of staged row j (the caller has seen the row land)
    auto patch_row = [&](int j) {
        uint8_t* const s = slot_of(j) + kRingPad - seg_x0;        // row position p lives at s[p]
        if (patch_left) *reinterpret_cast<uint32_t*>(s - 4) = (uint32_t)s[0] * 0x01010101u;
        if (patch_right) {
            const uint8_t e = s[W - 1];
#pragma unroll
            for (int k = 0; k < 11; ++k) s[W + k] = e;
        }
        stage::fence_generic_to_async();                          // the slot's next bulk copy overwrites these bytes
    };
    auto t3_put = [&](int j, const Tap3& v) { t3ring[(size_t)(j & (kT3Ring - 1)) * T + tid] = make_uint4(v.f[0], v.f[1], v.b[0], v.b[1]); };
    auto t3_get = [&](int j, Tap3& v) { const uint4 q = t3ring[(size_t)(j & (kT3Ring - 1)) * T + tid]; v.f[0] = q.x; v.f[1] = q.y; v.b[0] = q.z; v.b[1] = q.w; };
    // my 8 bytes of a picture row of the dst plane
    auto store8 = [&](int y, uint2 v) {
        uint8_t* const row = static_cast<uint8_t*>(t.plane) + (long long)y * t.pitch;
        if (vec_out) { *reinterpret_cast<uint2*>(row + x0) = v; return; }
#pragma unroll
        for (int b = 0; b < 8; ++b) if (b < npix) row[x0 + b] = (uint8_t)(((b < 4 ? v.x : v.y) >> (8 * (b & 3))) & 0xFFu);
    };

    uint2 pre[3] = {};                                  // cooperative mode: the row that goes into the ring next iteration
    if (seg_has_pixels) {
        if (bulk) {
            if (tid == 0) {
#pragma unroll
                for (int s = 0; s < kRing; ++s) stage::mbar_init(&mbar[s], 1);
                stage::fence_mbar_init();
            }
            __syncthreads();
            if (tid == 0)
                for (int j = 0; j <= kAhead && j < n; ++j) issue_row(j);
        } else {
            for (int j = 0; j < kAhead && j < n; ++j) { coop_load(j, pre); coop_store(j, pre); }
            if (kAhead < n) coop_load(kAhead, pre);
            __syncthreads();
        }
    }

    if (seg_has_pixels) {
        if (patch_left || patch_right)
            for (int j = 0; j < 3 && j < n; ++j) { await_row(j); patch_row(j); }
        __syncthreads();
    }

    // ---- running term M = B[r-1] + P[r] (+ leak); B[0] = 0 so M starts as P[1]. The only loop-carried registers. ----
    uint32_t M[kNumCost][4];

    // Stale costs of pool row `row` for my columns: what the previous pass of the frame left there.
    // (P[row] enters L[row-1] and L[row]; past the dependency cone of pool row `row - 1` no output of the frame can see
    // either, so those cells are not fetched at all: a warp stays until its FIRST column leaves the cone, this trims the
    // threads of its last rows that are already outside)
    auto stale_costs = [&](int row, uint32_t (&Pb)[kNumCost][2]) {
        StateRow in = state_row(t.in, row, x0, S);
        if (x0 >= t.cone - 3 * (row - 1)) in.p = nullptr;
#pragma unroll
        for (int i = 0; i < kNumCost; ++i) {
            const uint2 v = in.p != nullptr ? __ldg(reinterpret_cast<const uint2*>(in.p + i * in.stride)) : make_uint2(0u, 0u);
            Pb[i][0] = v.x; Pb[i][1] = v.y;
        }
    };
    // Pixel costs of a pair merged into stale costs for the thread that straddles the picture's right edge.
    auto straddle_costs = [&](const Taps& c, const Tap3& c3, const Taps& nx, const Tap3& n3, uint32_t (&Pb)[kNumCost][2]) {
        const uint32_t pixmask_lo = npix >= 4 ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (8 * (4 - npix)));
        const uint32_t pixmask_hi = npix <= 4 ? 0u : (0xFFFFFFFFu >> (8 * (8 - npix)));
        uint32_t Px[kNumCost][2];
        pair_costs(c, c3, nx, n3, Px);
#pragma unroll
        for (int i = 0; i < kNumCost; ++i) {
            Pb[i][0] = (Px[i][0] & pixmask_lo) | (Pb[i][0] & ~pixmask_lo);
            Pb[i][1] = (Px[i][1] & pixmask_hi) | (Pb[i][1] & ~pixmask_hi);
        }
    };

    {
        uint32_t Pb[kNumCost][2];
        if (npix < kCols || n < 2) stale_costs(1, Pb);
        if (npix > 0) {
            uint32_t wa[4], wb[4];
            Taps Ta, Tb;
            Tap3 ta, tb;
            await_row(0);
            window(0, wa);
            Ta.build(wa);
            tap3_row<kSat>(Ta, ta);
            t3_put(0, ta);
            // border row without a neighbour pair (reference GetFrame :380-391) and, for a one-pair-less plane, the kept row
            if (t.offset != 0 && !t.no_border) store8(0, make_uint2(wa[1], wa[2]));
            if (n == 1) {
                if (t.offset == 0 && !t.no_border) store8(t.height - 1, make_uint2(wa[1], wa[2]));
                if (t.copy_kept) store8(t.offset, make_uint2(wa[1], wa[2]));
            } else {
                await_row(1);
                window(1, wb);
                Tb.build(wb);
                tap3_row<kSat>(Tb, tb);
                t3_put(1, tb);
                if (npix == kCols) pair_costs(Ta, ta, Tb, tb, Pb); else straddle_costs(Ta, ta, Tb, tb, Pb);
            }
        }
#pragma unroll
        for (int i = 0; i < kNumCost; ++i) {
            M[i][0] = lanes_lo(Pb[i][0]) + leak(i); M[i][1] = lanes_hi(Pb[i][0]) + leak(i);
            M[i][2] = lanes_lo(Pb[i][1]) + leak(i); M[i][3] = lanes_hi(Pb[i][1]) + leak(i);
        }
    }
    // Threads without (all) pixel columns read the cost state the previous pass handed over. Those loads would open
    // every row and stall it for a DRAM round trip, so they run one row ahead: sp holds pool row r+1 at the top of row r.
#ifdef SN_HOST_EMULATION
    // the emulation has no warps: evaluate the predicate for the threads this one would share a warp with
    bool warp_full = true;
    for (int l = wfirst; l <= wlast; ++l) warp_full = warp_full && min(max(W - (seg_x0 + l * kCols), 0), kCols) == kCols;
#else
    const bool warp_full = __all_sync(0xFFFFFFFFu, npix == kCols);
#endif
    uint32_t sp[kNumCost][2];
    if (!warp_full) stale_costs(2, sp);
    // all blocks of a cluster run, with their halo barriers initialised, before the first DSMEM store
    if constexpr (kClustered) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int b = 0; b < 4; ++b) cl::halo_init(halo_bar(b >> 1, b & 1));
            cl::halo_fence_init();
        }
        cl::sync_all();
    }

    const uint32_t tkey = (uint32_t)min(t.thr_i + 1, 4095) * 0x00100010u;    // (thr+1) << 4 in both lanes
    // Dependency cone (sangnom_plan.h): at pool row r only columns < cone - 3r can still reach a picture sample of this
    // or a later pass of the frame. A warp whose first column lies beyond that has nothing left to do - it leaves, the
    // row barriers go on without it (warps with pixel columns stay to the last picture row by construction).
    const int r_last = min(R, (t.cone - 1 - (seg_x0 + wfirst * kCols)) / 3);

    // One pool row. kFull: every thread of the warp owns 8 pixel columns (no stale-state or masking code on the hot
    // path). kPair: row r+1 is a pair row (r + 1 <= n - 1), i.e. its costs come from pixels.
    // kExport: some thread of the warp hands this row's blurred costs to the next pass of the frame.
    auto row_step = [&](auto full, auto pairrow, auto exportrow, int r, const StateRow out) {
        constexpr bool kFull = decltype(full)::value, kPair = decltype(pairrow)::value, kExport = decltype(exportrow)::value;
        const bool pixels = kFull || npix > 0;
        // stage the ring
        if (seg_has_pixels) {
            if (bulk) {
                if (tid == 0 && r + kAhead < n) issue_row(r + kAhead);
            } else {
                if (r + kAhead - 1 < n) coop_store(r + kAhead - 1, pre);
                if (r + kAhead < n) coop_load(r + kAhead, pre);
            }
            if ((patch_left || patch_right) && r + 2 < n) { await_row(r + 2); patch_row(r + 2); }
        }
        uint32_t wb[4];                                  // K[r]: upper row of the pair whose costs are formed, lower row of the interpolation
        Taps Tb;
        Tap3 tb;

        // ---- P[r+1], L = M + P[r+1] -> shared row; M keeps P[r+1] until B[r] is known ----
        uint2* const Lrow = Lbase + ((size_t)(r & 1) * (T + 2) + 1 + tid) * kLEntry;      // my entry
        uint2 own_xy, own_zw;
        {
            uint32_t Pb[kNumCost][2];
            if constexpr (!kFull) {
#pragma unroll
                for (int i = 0; i < kNumCost; ++i) { Pb[i][0] = sp[i][0]; Pb[i][1] = sp[i][1]; }
            } else if constexpr (!kPair) {
                stale_costs(r + 1, Pb);
            }
            if (kPair && pixels) {
                uint32_t wc[4];
                Taps Tc;
                Tap3 tc;
                window(r, wb);
                Tb.build(wb);
                t3_get(r, tb);
                await_row(r + 1);
                window(r + 1, wc);
                Tc.build(wc);
                tap3_row<kSat>(Tc, tc);
                t3_put(r + 1, tc);
                if (kFull) pair_costs(Tb, tb, Tc, tc, Pb); else straddle_costs(Tb, tb, Tc, tc, Pb);
            }
            // last cost first: cost 0's own L words stay in registers for the start of phase B
#pragma unroll
            for (int ii = 0; ii < kNumCost; ++ii) {
                const int i = kNumCost - 1 - ii;
                const uint32_t P0 = lanes_lo(Pb[i][0]), P1 = lanes_hi(Pb[i][0]), P2 = lanes_lo(Pb[i][1]), P3 = lanes_hi(Pb[i][1]);
                const uint2 Lxy = make_uint2(M[i][0] + P0 - leak(i), M[i][1] + P1 - leak(i));
                const uint2 Lzw = make_uint2(M[i][2] + P2 - leak(i), M[i][3] + P3 - leak(i));
                Lrow[2 * i] = Lxy;
                Lrow[2 * i + 1] = Lzw;
                M[i][0] = P0; M[i][1] = P1; M[i][2] = P2; M[i][3] = P3;
                if (i == 0) { own_xy = Lxy; own_zw = Lzw; }
            }
        }
        // the two edge threads of the segment supply what lies beyond it: the clamp of the recursion at pool columns 0
        // and S-1 (reference :144-152), or - plane split over a cluster - the neighbour block's halo (DSMEM)
        // (the block right of mine is there as long as its first column is inside the cone; once it has left there is
        // nothing to send it and nothing to wait for)
        const bool right_block = kClustered && seg_last && !plane_last && 3 * r + seg_x0 + seg_cols < t.cone;
        [[maybe_unused]] const unsigned halo_parity = (unsigned)((r - 1) >> 1) & 1u;        // barrier [side][r & 1] completes its ((r-1)/2)-th phase at row r
        if (seg_first) {
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                const uint2 Lxy = Lrow[2 * i];
                if (plane_first) { const uint32_t e = (Lxy.x & 0xFFFFu) * 0x00010001u; Lrow[2 * i + 1 - kLEntry] = make_uint2(e, e); }
                else cl::store_remote_tx(&Lrow[2 * i + T * kLEntry], crank - 1, Lxy, halo_bar(1, r & 1));          // the left block's right halo
            }
        }
        if (seg_last) {
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                const uint2 Lzw = Lrow[2 * i + 1];
                if (plane_last) { const uint32_t e = (Lzw.y >> 16) * 0x00010001u; Lrow[2 * i + kLEntry] = make_uint2(e, e); }
                else if (right_block) cl::store_remote_tx(&Lrow[2 * i + 1 - T * kLEntry], crank + 1, Lzw, halo_bar(0, r & 1));   // the right block's left halo
            }
        }
        __syncthreads();
        if constexpr (kClustered) {                         // the one thread that reads a neighbour block's halo waits for it
            if (seg_first && !plane_first) cl::halo_wait(halo_bar(0, r & 1), halo_parity, kNumCost * 8u);
            if (right_block) cl::halo_wait(halo_bar(1, r & 1), halo_parity, kNumCost * 8u);
        }
        if constexpr (!kFull) stale_costs(r + 2, sp);       // next row's handed-over state: in flight during phase B

        // ---- per cost: 7-tap sum, key = (B << 4) | rank, M = P[r+1] + B (+ leak), min over the keys ----
        uint32_t kmin[4] = { tkey, tkey, tkey, tkey };
        uint32_t held[4];
        uint8_t* outp = out.p;
        // the four L words of a cost are fetched one cost ahead of their use, so that the shared-memory latency of
        // cost i+1 runs under the arithmetic of cost i
        uint2 nlh = Lrow[1 - kLEntry], noa = own_xy, nob = own_zw, nrh = Lrow[kLEntry];
#pragma unroll
        for (int i = 0; i < kNumCost; ++i) {
            const uint2 lh = nlh;                           // (l-4,l-3) (l-2,l-1)
            const uint2 oa = noa;                           // (l0,l1) (l2,l3)
            const uint2 ob = nob;                           // (l4,l5) (l6,l7)
            const uint2 rh = nrh;                           // (l8,l9) (l10,l11)
            if (i + 1 < kNumCost) {
                nlh = Lrow[2 * (i + 1) + 1 - kLEntry]; noa = Lrow[2 * (i + 1)]; nob = Lrow[2 * (i + 1) + 1]; nrh = Lrow[2 * (i + 1) + kLEntry];
            }
            const uint32_t Wm2 = lh.x, Wm1 = lh.y, W0 = oa.x, W1 = oa.y, W2 = ob.x, W3 = ob.y, W4 = rh.x, W5 = rh.y;
            // Z_k = W[k-1]+W[k]+W[k+1] (even / odd triples), X_k = Z_k + W[k-2];
            // H7_k = Z_k + (X_k.hi, X_{k+1}.lo)  -> lanes (sum l[2k-3..2k+3], sum l[2k-2..2k+4])
            const uint32_t Z0 = Wm1 + W0 + W1, Z1 = W0 + W1 + W2, Z2 = W1 + W2 + W3, Z3 = W2 + W3 + W4, Z4 = W3 + W4 + W5;
            const uint32_t X0 = Z0 + Wm2, X1 = Z1 + Wm1, X2 = Z2 + W0, X3 = Z3 + W1, X4 = Z4 + W2;
            // (B << 4) | rank in both lanes, B = the blurred cost narrowed to 8 bits: wrapped (:152), or clamped for the SSE2 flavour
            auto narrow_key = [&](uint32_t sum) -> uint32_t {
                if constexpr (kSat) return __vminu2(sum & 0xFFF0FFF0u, 0x0FF00FF0u) | rank2(i);
                else return (sum & keymask) | rank2(i);
            };
            uint32_t key[4];
            key[0] = narrow_key(Z0 + __funnelshift_r(X0, X1, 16));
            key[1] = narrow_key(Z1 + __funnelshift_r(X1, X2, 16));
            key[2] = narrow_key(Z2 + __funnelshift_r(X2, X3, 16));
            key[3] = narrow_key(Z3 + __funnelshift_r(X3, X4, 16));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                M[i][q] += key[q] >> 4;                                     // LEA.HI: P[r+1] + B[r] + leak(i)
                if (i & 1) kmin[q] = __vimin3_u16x2(kmin[q], held[q], key[q]);
                else if (i == kNumCost - 1) kmin[q] = __vminu2(kmin[q], key[q]);
                else held[q] = key[q];
            }
            // hand the blurred row to the next pass of this frame
            if (kExport) {
                if (out.p != nullptr) *reinterpret_cast<uint2*>(outp) = make_uint2(pack4(key[0] >> 4, key[1] >> 4), pack4(key[2] >> 4, key[3] >> 4));
                outp += out.stride;
            }
        }

        // ---- interpolate the picture row between K[r-1] and K[r] ----
        if (pixels && (kPair || r == n - 1)) {
            uint32_t wa[4];
            Taps Ta;
            Tap3 ta;
            window(r - 1, wa);
            Ta.build(wa);
            t3_get(r - 1, ta);
            if (!kPair) { window(r, wb); Tb.build(wb); t3_get(r, tb); }
            const uint2 px = interpolate8(Ta, ta, Tb, tb, kmin);
            const int y = t.offset + 2 * (r - 1);
            store8(y + 1, px);
            if (t.copy_kept) store8(y, make_uint2(wa[1], wa[2]));
            if (!kPair) {                                                   // K[r] is the last kept row
                if (t.offset == 0 && !t.no_border) store8(t.height - 1, make_uint2(wb[1], wb[2]));
                if (t.copy_kept) store8(y + 2, make_uint2(wb[1], wb[2]));
            }
        }
    };
    // The rows my warp exports are two ranges known up front (region B: rows b_r0..b_r1, all columns; region A: rows
    // 1..a_rows for the warps that reach past a_x0), so the per-row question is two register compares; the task's
    // region description is read from shared memory only for rows that are exported.
    int ex_b0 = 1, ex_b1 = 0, ex_a1 = 0;
    if (t.out.b != nullptr) { ex_b0 = t.out.b_r0; ex_b1 = t.out.b_r1; }
    {
        const bool mine_a = t.out.a != nullptr && x0 >= t.out.a_x0;
#ifdef SN_HOST_EMULATION
        bool warp_a = false;
        for (int l = wfirst; l <= wlast; ++l) warp_a = warp_a || (t.out.a != nullptr && seg_x0 + l * kCols >= t.out.a_x0);
#else
        const bool warp_a = __any_sync(0xFFFFFFFFu, mine_a);
#endif
        if (warp_a) ex_a1 = t.out.a_rows;
    }
    // does my warp export pool row r? (warp-uniform, so the variants of a row keep warps whole)
    auto export_row = [&](int r, StateRow& out) -> bool {
        out = StateRow{ nullptr, 0 };
        if (!((r >= ex_b0 && r <= ex_b1) || r <= ex_a1)) return false;
        if (x0 < t.export_cone - 3 * r) out = state_row(t.out, r, x0, S);         // beyond it nothing downstream reads the row
        return true;
    };
    auto sweep = [&](auto full) {
        int r = 1;
        StateRow out;
        for (; r <= n - 2 && r <= r_last; ++r) {                            // rows whose lower neighbour row is a pair row
            if (export_row(r, out)) row_step(full, std::true_type{}, std::true_type{}, r, out);
            else row_step(full, std::true_type{}, std::false_type{}, r, out);
        }
        for (; r <= r_last; ++r) {                                          // the last picture row and rows swept for the next pass only
            if (export_row(r, out)) row_step(full, std::false_type{}, std::true_type{}, r, out);
            else row_step(full, std::false_type{}, std::false_type{}, r, out);
        }
    };
    // warp-uniform choice, so that a warp never splits over the two copies of the row barrier
    if (warp_full) sweep(std::true_type{}); else sweep(std::false_type{});
#ifdef SN_HOST_EMULATION
    if (r_last < R) {                                                       // left before the last row: out of the barriers
        emul::bar->arrive_and_drop();
        if (kClustered) emul::cluster_bar->arrive_and_drop();
    }
#endif
}

}  // namespace u8k
}  // namespace sn
