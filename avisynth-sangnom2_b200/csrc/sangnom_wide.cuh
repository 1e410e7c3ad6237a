// 16-bit and fp32 flavours of the fused row sweep (one sample per 32-bit lane), built on the same skeleton as the
// 8-bit kernel (sangnom_u8.cuh): kept rows staged into a shared-memory ring by bulk async copies, one barrier per
// pool row, only the running term M loop-carried in registers.
//
// One thread owns 4 adjacent pool columns. Per pool row r:
//   phase A  kept row K[r+1] arrives in the ring; the thread takes its 12-sample window (columns x0-4 .. x0+7),
//            forms the two 3-tap values of every own pixel of that row (once per row: they serve two pairs and the
//            interpolation) and parks them in a small shared ring, forms the nine raw costs P[r+1] of the pair
//            (K[r], K[r+1]) and publishes L = M + P[r+1] to a double-buffered shared row. The L row is entry-major:
//            the nine 16-byte vectors of a thread are 144 bytes apart from the next thread's, so every access is a
//            conflict-free 128-bit one and every per-cost offset an immediate;
//   barrier  one block barrier per row; when the plane is split over the blocks of a cluster, the two edge threads
//            of a block also exchange their halo with the neighbour blocks (sangnom_cluster.cuh) - no cluster barrier;
//   phase B  per cost: three 128-bit reads (left neighbour, own, right neighbour), the 7-tap sum (integers: sliding,
//            1.5 adds per column; fp32: the reference's left-to-right order, each add rounded), /16, narrow to T,
//            min key (integers: (B << 4) | rank with the threshold as a tenth key; fp32: strict '<' in tie order),
//            M = P[r+1] + B; then the interpolated picture row between K[r-1] and K[r]: the winning direction's two
//            operands are fetched from the ring by index (no select chains, no pixel windows kept in registers).
// State carried in registers: M (36). Warps whose columns have left the dependency cone of every remaining output
// retire (sangnom_plan.h).
//
// Reference semantics: /root/reference/src/SangNom2.cpp :25-34 (edge replication), :60-72 (3-tap), :74-124 (costs),
// :126-159 (recursive cost sum, /16, narrowing), :161-257 (min, threshold, tie order, mean), GetFrame :361-391.
#pragma once
#include "sangnom_arith.cuh"
#include "sangnom_cluster.cuh"
#include "sangnom_stage.cuh"

#include <type_traits>

#ifndef SN_DYNAMIC_SMEM
#define SN_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace sn {
namespace wide {

constexpr int kCols = 4;                 // pool columns per thread
constexpr int kWin = 12;                 // window: samples x0-4 .. x0+7
constexpr int kLEntry = kNumCost;        // uint4 words per thread entry of the shared L rows (144 B)
constexpr int kRing = 5;                 // kept-row ring slots: rows r-1 .. r+1 in use, row r+2 in flight, and one more so that the slot
                                         // a copy lands in was last read two barriers ago (row r-3). Not a power of two on purpose: at
                                         // 4 bytes per sample a 960-column block must stay below half an SM's shared memory.
constexpr int kAhead = 2;                // rows staged ahead of the row being consumed
constexpr int kT3Ring = 3;               // 3-tap rows r-1, r, r+1
constexpr int kRingPadBytes = 16;        // halo on each side of a staged row segment (>= 3 samples of every type)

// d + 3 for the winning rank's horizontal tap d (upper row: +d, lower row: -d), one nibble per rank 0..7; rank 8 = 0
// rank: 0 vertical, 1 (b1,b2), 2 (f1,f2), 3 (+1,-1), 4 (-1,+1), 5 (+2,-2), 6 (-2,+2), 7 (+3,-3), 8 (-3,+3)
constexpr unsigned kTapTable = 0x61524333u;

__device__ __forceinline__ int tap_index(int rank)
{
#ifdef SN_HOST_EMULATION
    return rank >= 8 ? 0 : (int)((kTapTable >> (4 * rank)) & 15u);
#else
    return (int)(__funnelshift_rc(kTapTable, 0u, 4 * rank) & 15u);
#endif
}

// cost buffer visited k-th in phase B: fp32 goes through the reference's tie order, integers 0..8
template <bool kTieOrder> __device__ __forceinline__ constexpr int visit_cost(int k)
{
    constexpr int order[kNumCost] = { 4, 5, 3, 6, 2, 7, 1, 8, 0 };
    return kTieOrder ? order[k] : k;
}

template <typename T> struct Vec4;       // the 4 samples of a thread as one vector
template <> struct Vec4<uint16_t> { using type = uint2; };
template <> struct Vec4<float> { using type = float4; };

__device__ __forceinline__ void unpack4(const uint2 v, int (&o)[4])
{
    o[0] = (int)(v.x & 0xFFFFu); o[1] = (int)(v.x >> 16); o[2] = (int)(v.y & 0xFFFFu); o[3] = (int)(v.y >> 16);
}
__device__ __forceinline__ void unpack4(const float4 v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ uint2 pack4(const int (&v)[4], uint16_t)
{
    return make_uint2(((uint32_t)v[0] & 0xFFFFu) | ((uint32_t)v[1] << 16), ((uint32_t)v[2] & 0xFFFFu) | ((uint32_t)v[3] << 16));
}
__device__ __forceinline__ float4 pack4(const float (&v)[4], float) { return make_float4(v[0], v[1], v[2], v[3]); }

__device__ __forceinline__ uint4 as_words(const int (&v)[4]) { return make_uint4((uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3]); }
__device__ __forceinline__ uint4 as_words(const float (&v)[4])
{
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
}
__device__ __forceinline__ void from_words(const uint4 q, int (&v)[4]) { v[0] = (int)q.x; v[1] = (int)q.y; v[2] = (int)q.z; v[3] = (int)q.w; }
__device__ __forceinline__ void from_words(const uint4 q, float (&v)[4])
{
    v[0] = __uint_as_float(q.x); v[1] = __uint_as_float(q.y); v[2] = __uint_as_float(q.z); v[3] = __uint_as_float(q.w);
}
__device__ __forceinline__ uint32_t word_of(int v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t word_of(float v) { return __float_as_uint(v); }

// Where the cost state of one pool row lives for this thread's 4 columns: samples of buffer i at p + i * stride.
// p == nullptr: outside the handed-over regions (reads as the zero-filled pool, nothing to write).
template <typename T> struct StateRow { T* p; size_t stride; };

template <typename T>
__device__ __forceinline__ StateRow<T> state_row(const CostState& s, int r, int x0, int S)
{
    if (s.b != nullptr && r >= s.b_r0 && r <= s.b_r1) {
        const int nb = s.b_r1 - s.b_r0 + 1;
        return StateRow<T>{ static_cast<T*>(s.b) + (size_t)(r - s.b_r0) * S + x0, (size_t)nb * S };
    }
    if (s.a != nullptr && x0 >= s.a_x0 && r >= 1 && r <= s.a_rows) {
        const int wa = S - s.a_x0;
        return StateRow<T>{ static_cast<T*>(s.a) + (size_t)r * wa + (x0 - s.a_x0), (size_t)(s.a_rows + 1) * wa };
    }
    return StateRow<T>{ nullptr, 0 };
}

// 4 samples at row positions p0 .. p0+3 with zero outside [0, W); `fast`: aligned vector loads are legal
template <typename T>
__device__ __forceinline__ typename Vec4<T>::type load4_guarded(const T* __restrict__ row, int p0, int W, bool fast)
{
    using V = typename Vec4<T>::type;
    if (fast && p0 >= 0 && p0 + 4 <= W) return *reinterpret_cast<const V*>(row + p0);
    T v[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) { const int p = p0 + b; v[b] = (p >= 0 && p < W) ? row[p] : T(0); }
    V out;
    memcpy(&out, v, sizeof out);
    return out;
}

// Shared memory of one block (seg_cols pool columns, T = seg_cols / 4 threads):
//   L     [2 parities][T + 2 entries][9 costs] uint4: the vertical sums of a thread's 4 columns; entries 0 and T+1 hold
//         the neighbour segment's edge or the clamp
//   ring  [kRing][ring_stride]   staged kept rows, row position p at byte offset (p - seg_x0) * sizeof(T) + kRingPadBytes
//   t3    [kT3Ring][2][seg_cols] the two 3-tap values of every pixel of a kept row, as samples (f, then b)
//   mbar  [kRing]                one mbarrier per ring slot
//   halo  [2 sides][2 parities]  barriers the neighbour blocks' halo stores complete on (cluster launches)
//   task                         this block's PlaneTask
template <typename T> inline __host__ __device__ int ring_stride(int seg_cols) { return (seg_cols * (int)sizeof(T) + 2 * kRingPadBytes + 15) & ~15; }
inline __host__ __device__ size_t l_bytes(int seg_cols) { return (size_t)2 * (seg_cols / kCols + 2) * kLEntry * sizeof(uint4); }
template <typename T> inline size_t smem_bytes(int seg_cols)
{
    return l_bytes(seg_cols) + (size_t)kRing * ring_stride<T>(seg_cols) + (size_t)kT3Ring * 2 * seg_cols * sizeof(T) +
           kRing * sizeof(stage::Mbar) + 4 * 16 /* halo barriers */ + ((sizeof(PlaneTask) + 15) & ~(size_t)15);
}

// kSpare: the launch brings spare threads for planes narrower than the pool (see the thread -> column map below).
template <typename T, int kMaxThreads, int kMinBlocks, bool kClustered, bool kSat = false, bool kSpare = false>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks)
sangnom_wide_row_sweep(const PlaneTask* __restrict__ tasks, LaunchGeometry g, int seg_cols)
{
    using I = typename Flavour<T>::I;
    using V = typename Vec4<T>::type;
    constexpr bool kFloat = Flavour<T>::kFloat;
    constexpr int kPadS = kRingPadBytes / (int)sizeof(T);           // ring halo in samples
    SN_DYNAMIC_SMEM(smem_raw);

    const unsigned G = kClustered ? cl::size() : 1u;      // kClustered = false: one block per plane, no cluster code at all
    const unsigned crank = kClustered ? cl::rank() : 0u;
    const int S = g.S;
    const int Tn = seg_cols / kCols;                                // working threads of the block
    uint4* const Lbase = reinterpret_cast<uint4*>(smem_raw);
    const int rstride = ring_stride<T>(seg_cols);
    unsigned char* const ring = smem_raw + l_bytes(seg_cols);
    T* const t3ring = reinterpret_cast<T*>(ring + (size_t)kRing * rstride);
    stage::Mbar* const mbar = reinterpret_cast<stage::Mbar*>(t3ring + (size_t)kT3Ring * 2 * seg_cols);
    // halo barriers, 16 bytes apart: [side 0 = left, 1 = right][row parity]
    unsigned char* const halo_raw = reinterpret_cast<unsigned char*>(mbar + kRing);
    auto halo_bar = [&](int side, int par) -> cl::HaloBar* { return reinterpret_cast<cl::HaloBar*>(halo_raw + (side * 2 + par) * 16); };
    // the task lives in shared memory: its rarely used fields are re-read where needed instead of occupying registers
    {
        uint32_t* const dst = reinterpret_cast<uint32_t*>(halo_raw + 4 * 16);
        const uint32_t* const from = reinterpret_cast<const uint32_t*>(tasks + blockIdx.x / G);
        for (unsigned k = threadIdx.x; k < sizeof(PlaneTask) / 4; k += blockDim.x) dst[k] = from[k];
        __syncthreads();
    }
    const PlaneTask& t = *reinterpret_cast<const PlaneTask*>(halo_raw + 4 * 16);

    const int W = t.width, n = t.kept_rows, R = t.sweep_rows;
    const int seg_x0 = (int)crank * seg_cols;
    // Thread -> column map: as in the 8-bit kernel, a plane narrower than the pool gets spare threads and the first
    // `shift` threads stay idle so that the last pixel thread ends a warp (no warp mixes pixel and state-only lanes).
    const int hw = (int)threadIdx.x;
    const int Tpx = (min(max(W - seg_x0, 0), seg_cols) + kCols - 1) / kCols;       // threads of this segment that carry pixels
    int shift = (kSpare && !kClustered && Tpx > 0 && Tpx < Tn && (Tpx & 31) != 0) ? 32 - (Tpx & 31) : 0;
    if (Tn + shift > (int)blockDim.x) shift = 0;
    const int tid = kSpare ? hw - shift : hw;                       // working thread index (block-uniform offset)
    if (kSpare && (tid < 0 || tid >= Tn)) {                         // spare thread: nothing to do, not even the barriers
#ifdef SN_HOST_EMULATION
        emul::bar->arrive_and_drop();
#endif
        return;
    }
    const int wfirst = max((hw & ~31) - shift, 0);                  // first working thread of my warp
#ifdef SN_HOST_EMULATION
    const int wlast = min((hw & ~31) + 31 - shift, Tn - 1);
#endif
    const int lx = tid * kCols;                                     // column inside the segment
    const int x0 = seg_x0 + lx;                                     // pool column
    const bool plane_first = x0 == 0, plane_last = x0 + kCols == S;
    const bool seg_first = tid == 0, seg_last = tid == Tn - 1;
    const T* const src = static_cast<const T*>(t.src);
    const long long src_pitch = t.src_pitch;                        // samples
    const int wpad = (W * (int)sizeof(T) + 15) / 16 * (16 / (int)sizeof(T));       // W rounded up to 16 bytes, in samples
    const int npix = min(max(W - x0, 0), kCols);                    // how many of my columns carry pixels
    const bool edge = npix > 0 && (x0 == 0 || x0 + 7 > W - 1);     // a picture edge inside my window
    const bool vec_out = ((reinterpret_cast<uintptr_t>(t.plane) | (uintptr_t)(t.pitch * (long long)sizeof(T))) & (sizeof(V) - 1)) == 0 && npix == kCols;

    // ---- staging of kept rows: positions [lo, hi) of every kept row go to ring offset (position - seg_x0 + kPadS) ----
    const int lo = max(seg_x0 - kPadS, 0), hi = min(seg_x0 + seg_cols + kPadS, wpad);
    const bool seg_has_pixels = seg_x0 < W;
    const bool bulk = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(src_pitch * (long long)sizeof(T)) | (uintptr_t)(seg_cols * (int)sizeof(T))) & 15) == 0 && src_pitch >= wpad;
    const bool fastv = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(src_pitch * (long long)sizeof(T))) & (sizeof(V) - 1)) == 0;
    auto kept_row = [&](int j) -> const T* { return src + (long long)j * src_pitch; };
    // Ring slots are named by slot index (row j lives in slot j mod kRing); the sweep carries the slot of row r+1 and
    // the parity of its mbarrier phase along instead of dividing.
    auto slot_ptr = [&](int slot) -> T* { return reinterpret_cast<T*>(ring + (size_t)slot * rstride); };
    auto wrap = [](int slot) -> int { return slot >= kRing ? slot - kRing : (slot < 0 ? slot + kRing : slot); };
    auto issue_row = [&](int j, int slot) {
        stage::bulk_load(slot_ptr(slot) + (lo - seg_x0 + kPadS), kept_row(j) + lo, (unsigned)((hi - lo) * (int)sizeof(T)), &mbar[slot]);
    };
    // cooperative mode (unaligned sources): my own 4 samples of a row, edge threads also the neighbour segments' halo
    auto coop_load = [&](int j, V (&v)[3]) {
        const T* row = kept_row(j);
        v[0] = load4_guarded<T>(row, x0, W, fastv);
        if (kClustered && seg_first) v[1] = load4_guarded<T>(row, x0 - 4, W, fastv);
        if (kClustered && seg_last) v[2] = load4_guarded<T>(row, x0 + 4, W, fastv);
    };
    auto coop_store = [&](int slot, const V (&v)[3]) {
        T* s = slot_ptr(slot) + kPadS + lx;
        *reinterpret_cast<V*>(s) = v[0];
        if (kClustered && seg_first) *reinterpret_cast<V*>(s - 4) = v[1];
        if (kClustered && seg_last) *reinterpret_cast<V*>(s + 4) = v[2];
    };
    auto await_row = [&](int slot, unsigned parity) { if (bulk) stage::mbar_wait(&mbar[slot], parity); };
    // Picture edges (reference loadPixel :25-34: sample 0 and sample W-1 replicated outwards) are made in the staged row
    // itself, once, by the threads whose window holds an edge, right after they have seen the row land and before they
    // read it: every position a thread ever reads of a kept row (window x0-4 .. x0+7, interpolation operands x0-3 ..
    // x0+6) lies inside its own window, so each edge thread fixes what it will read and nobody waits for anybody; where
    // two right-edge threads overlap (W not a multiple of 4) they write the same value. After that neither the windows
    // nor the indexed operand fetch of the interpolation know about edges. The positions left of 0 are never written by
    // a bulk copy; those right of W-1 only when W is not a multiple of 16 bytes (then the next copy into the slot is
    // ordered behind these stores by a proxy fence).
    auto patch_edges = [&](int slot) {
        T* const s = slot_ptr(slot) + kPadS + lx;                   // my sample x0
        if (x0 == 0) { const T v = s[0]; s[-4] = v; s[-3] = v; s[-2] = v; s[-1] = v; }
        const int last = W - 1 - x0;                                // my index of the last picture sample (>= 0)
        if (last < 7) {
            const T v = s[last];
#pragma unroll
            for (int e = 1; e < 8; ++e) if (e > last) s[e] = v;
            if (wpad != W) stage::fence_generic_to_async();
        }
    };
    // my window of the kept row in `slot` (samples x0-4 .. x0+7) out of the ring (edges: see patch_edges)
    auto window = [&](int slot, I (&w)[kWin]) {
        const T* s = slot_ptr(slot) + kPadS + lx;
        I a[4], b[4], c[4];
        unpack4(*reinterpret_cast<const V*>(s - 4), a);
        unpack4(*reinterpret_cast<const V*>(s), b);
        unpack4(*reinterpret_cast<const V*>(s + 4), c);
#pragma unroll
        for (int e = 0; e < 4; ++e) { w[e] = a[e]; w[4 + e] = b[e]; w[8 + e] = c[e]; }
    };
    // 3-tap values of a kept row for my 4 columns: f = T3(x-1, x, x+1), b = T3(x+1, x, x-1). As the upper row of a
    // pair they are (f1, b1), as the lower row (f2, b2) = (b, f) (reference :103-106).
    struct Tap3 { I f[4], b[4]; };
    auto tap3_row = [&](const I (&w)[kWin], Tap3& o) {
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
            o.f[c] = tap3<T, I, kSat>(w[c + 3], w[c + 4], w[c + 5]);
            o.b[c] = tap3<T, I, kSat>(w[c + 5], w[c + 4], w[c + 3]);
        }
    };
    // slot of row j in the 3-tap ring: j mod 3 is carried along the sweep (ph3 = r mod 3), never divided out
    auto t3_row = [&](int slot) -> T* { return t3ring + (size_t)slot * 2 * seg_cols + lx; };
    auto t3_put = [&](int slot, const Tap3& v) {
        T* p = t3_row(slot);
        *reinterpret_cast<V*>(p) = pack4(v.f, T());
        *reinterpret_cast<V*>(p + seg_cols) = pack4(v.b, T());
    };
    auto t3_get = [&](int slot, Tap3& v) {
        const T* p = t3_row(slot);
        unpack4(*reinterpret_cast<const V*>(p), v.f);
        unpack4(*reinterpret_cast<const V*>(p + seg_cols), v.b);
    };
    // my 4 samples of a picture row of the dst plane
    auto store4 = [&](int y, const V v) {
        T* const row = static_cast<T*>(t.plane) + (long long)y * t.pitch;
        if (vec_out) { *reinterpret_cast<V*>(row + x0) = v; return; }
        T s[4];
        memcpy(s, &v, sizeof v);
#pragma unroll
        for (int b = 0; b < 4; ++b) if (b < npix) row[x0 + b] = s[b];
    };
    auto own4 = [&](int slot) -> V { return *reinterpret_cast<const V*>(slot_ptr(slot) + kPadS + lx); };

    V pre[3] = {};                                      // cooperative mode: the row that goes into the ring next iteration
    if (seg_has_pixels) {
        if (bulk) {
            if (tid == 0) {
#pragma unroll
                for (int s = 0; s < kRing; ++s) stage::mbar_init(&mbar[s], 1);
                stage::fence_mbar_init();
            }
            __syncthreads();
            if (tid == 0)
                for (int j = 0; j <= kAhead && j < n; ++j) issue_row(j, j);
        } else {
            // rows 0 .. kAhead into the ring, row kAhead+1 into registers: a row is stored one iteration before its
            // first use, so that the row barrier in between orders the store before the neighbours' reads
            for (int j = 0; j <= kAhead && j < n; ++j) { coop_load(j, pre); coop_store(j, pre); }
            if (kAhead + 1 < n) coop_load(kAhead + 1, pre);
            __syncthreads();
        }
    }

    // ---- running term M = B[r-1] + P[r]; B[0] = 0 so M starts as P[1]. The only loop-carried registers. ----
    I M[kNumCost][kCols];

    // Stale costs of pool row `row` for my columns: what the previous pass of the frame left there.
    auto stale_costs = [&](int row, I (&Pb)[kNumCost][kCols]) {
        const StateRow<T> in = state_row<T>(t.in, row, x0, S);
#pragma unroll
        for (int i = 0; i < kNumCost; ++i) {
            if (in.p != nullptr) unpack4(__ldg(reinterpret_cast<const V*>(in.p + i * in.stride)), Pb[i]);
            else { Pb[i][0] = Pb[i][1] = Pb[i][2] = Pb[i][3] = I(0); }
        }
    };
    // The same cells pulled towards L1 a row ahead of their use: threads without (all) pixel columns read the previous
    // pass's cost state every row, and holding a row of it in registers across phase B (as the 8-bit kernel does) costs
    // 36 registers this kernel does not have - it spilled. A prefetch costs none.
    auto prefetch_stale = [&](int row) {
#ifndef SN_HOST_EMULATION
        const StateRow<T> in = state_row<T>(t.in, row, x0, S);
        if (in.p != nullptr) {
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(in.p + i * in.stride));
        }
#else
        (void)row;
#endif
    };
    // Nine raw costs of the pair (upper row window u / 3-tap u3, lower row l / l3) for my columns; columns without
    // pixels keep what Pb holds (the straddling thread's stale costs).
    auto pair_costs = [&](auto full, const I (&u)[kWin], const Tap3& u3, const I (&l)[kWin], const Tap3& l3, I (&Pb)[kNumCost][kCols]) {
        constexpr bool kAll = decltype(full)::value;
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
            if (!kAll && c >= npix) continue;
            const int q = c + 4;
            Pb[0][c] = absdiff(u[q - 3], l[q + 3]);
            Pb[1][c] = absdiff(u[q - 2], l[q + 2]);
            Pb[2][c] = absdiff(u[q - 1], l[q + 1]);
            Pb[3][c] = absdiff(u3.f[c], l3.b[c]);           // |f1 - f2|
            Pb[4][c] = absdiff(u[q], l[q]);
            Pb[5][c] = absdiff(u3.b[c], l3.f[c]);           // |b1 - b2|
            Pb[6][c] = absdiff(u[q + 1], l[q - 1]);
            Pb[7][c] = absdiff(u[q + 2], l[q - 2]);
            Pb[8][c] = absdiff(u[q + 3], l[q - 3]);
        }
    };

    {
        I Pb[kNumCost][kCols];
        if (npix < kCols || n < 2) stale_costs(1, Pb);
        if (npix > 0) {
            I wa[kWin], wb[kWin];
            Tap3 ta, tb;
            await_row(0, 0u);
            if (edge) patch_edges(0);
            window(0, wa);
            tap3_row(wa, ta);
            t3_put(0, ta);
            const I own[4] = { wa[4], wa[5], wa[6], wa[7] };
            // border row without a neighbour pair (reference GetFrame :380-391) and, for a one-pair-less plane, the kept row
            if (t.offset != 0 && !t.no_border) store4(0, pack4(own, T()));
            if (n == 1) {
                if (t.offset == 0 && !t.no_border) store4(t.height - 1, pack4(own, T()));
                if (t.copy_kept) store4(t.offset, pack4(own, T()));
            } else {
                await_row(1, 0u);
                if (edge) patch_edges(1);
                window(1, wb);
                tap3_row(wb, tb);
                t3_put(1, tb);
                if (npix == kCols) pair_costs(std::true_type{}, wa, ta, wb, tb, Pb); else pair_costs(std::false_type{}, wa, ta, wb, tb, Pb);
            }
        }
#pragma unroll
        for (int i = 0; i < kNumCost; ++i)
#pragma unroll
            for (int c = 0; c < kCols; ++c) M[i][c] = Pb[i][c];
    }
    // Threads without (all) pixel columns read the cost state the previous pass handed over every row (prefetched a row
    // ahead, see prefetch_stale).
#ifdef SN_HOST_EMULATION
    bool warp_full = true;
    for (int l = wfirst; l <= wlast; ++l) warp_full = warp_full && min(max(W - (seg_x0 + l * kCols), 0), kCols) == kCols;
#else
    const bool warp_full = __all_sync(0xFFFFFFFFu, npix == kCols);
#endif
    if (!warp_full) prefetch_stale(2);
    // all blocks of a cluster run, with their halo barriers initialised, before the first DSMEM store
    if constexpr (kClustered) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int b = 0; b < 4; ++b) cl::halo_init(halo_bar(b >> 1, b & 1));
            cl::halo_fence_init();
        }
        cl::sync_all();
    }

    const int tkey = (int)min((long long)t.thr_i + 1, 0x7FFFFFFLL) << 4;      // (thr+1) << 4: "every cost above the threshold"
    const unsigned keymask = g.key_mask;                            // kMask << 4, kept in a register so that (sum & mask) | rank is one LOP3
    const float thr_f = t.thr_f;
    // dependency cone (sangnom_plan.h): the last pool row my warp still has anything to do at
    const int r_last = min(R, (t.cone - 1 - (seg_x0 + wfirst * kCols)) / 3);

    // One pool row. kFull: every thread of the warp owns 4 pixel columns. kPair: pool row r+1 is a pair row
    // (r + 1 <= n - 1). kExport: some thread of the warp hands this row's blurred costs to the next pass.
    // ph3 = r mod 3 (3-tap ring slot of kept row r); s1 / par1 = ring slot of kept row r+1 and the parity of its
    // mbarrier phase.
    auto row_step = [&](auto full, auto pairrow, auto exportrow, int r, int ph3, int s1, unsigned par1, const StateRow<T> out) {
        constexpr bool kFull = decltype(full)::value, kPair = decltype(pairrow)::value, kExport = decltype(exportrow)::value;
        const bool pixels = kFull || npix > 0;
        const int ph3_next = ph3 == 2 ? 0 : ph3 + 1, ph3_prev = ph3 == 0 ? 2 : ph3 - 1;
        const int s0 = wrap(s1 - 1), sm1 = wrap(s1 - 2);                    // slots of kept rows r and r-1
        // stage the ring: row r+2 goes where row r-3 was (last read in iteration r-2, two barriers ago)
        if (seg_has_pixels) {
            if (bulk) {
                if (tid == 0 && r + kAhead < n) issue_row(r + kAhead, wrap(s1 + 1));
            } else {
                if (r + kAhead < n) coop_store(wrap(s1 + 1), pre);
                if (r + kAhead + 1 < n) coop_load(r + kAhead + 1, pre);
            }
        }

        // ---- P[r+1], L = M + P[r+1] -> shared row; M keeps P[r+1] until B[r] is known ----
        uint4* const Lrow = Lbase + ((size_t)(r & 1) * (Tn + 2) + 1 + tid) * kLEntry;      // my entry
        I own0[4];                                                                        // my own L values of the first cost visited in phase B
        {
            I Pb[kNumCost][kCols];
            if constexpr (!kFull || !kPair) stale_costs(r + 1, Pb);
            if (kPair && pixels) {
                I wb[kWin], wc[kWin];
                Tap3 tb, tc;
                window(s0, wb);
                t3_get(ph3, tb);
                await_row(s1, par1);
                if (edge) patch_edges(s1);
                window(s1, wc);
                tap3_row(wc, tc);
                t3_put(ph3_next, tc);
                if (kFull) pair_costs(std::true_type{}, wb, tb, wc, tc, Pb); else pair_costs(std::false_type{}, wb, tb, wc, tc, Pb);
            }
#pragma unroll
            for (int ii = 0; ii < kNumCost; ++ii) {
                const int i = kNumCost - 1 - ii;
                I L[4];
#pragma unroll
                for (int c = 0; c < kCols; ++c) { L[c] = add2(M[i][c], Pb[i][c]); M[i][c] = Pb[i][c]; }
                Lrow[i] = as_words(L);
                if (i == 0) { own0[0] = L[0]; own0[1] = L[1]; own0[2] = L[2]; own0[3] = L[3]; }
            }
        }
        // the two edge threads of the segment supply what lies beyond it: the clamp of the recursion at pool columns 0
        // and S-1 (reference :144-152 clamps at the pool stride), or - plane split over a cluster - the neighbour
        // block's halo (DSMEM). The neighbour reads columns 1..3 of its left halo entry and 0..2 of its right one.
        // (the block right of mine is there as long as its first column is inside the cone; once it has left there is
        // nothing to send it and nothing to wait for)
        const bool right_block = kClustered && seg_last && !plane_last && 3 * r + seg_x0 + seg_cols < t.cone;
        [[maybe_unused]] const unsigned halo_parity = (unsigned)((r - 1) >> 1) & 1u;        // barrier [side][r & 1] completes its ((r-1)/2)-th phase at row r
        if (seg_first) {
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                const uint4 Lv = Lrow[i];
                if (plane_first) Lrow[i - kLEntry] = make_uint4(Lv.x, Lv.x, Lv.x, Lv.x);
                else cl::store_remote_tx(&Lrow[i + Tn * kLEntry], crank - 1, Lv, halo_bar(1, r & 1));      // the left block's right halo
            }
        }
        if (seg_last) {
#pragma unroll
            for (int i = 0; i < kNumCost; ++i) {
                const uint4 Lv = Lrow[i];
                if (plane_last) Lrow[i + kLEntry] = make_uint4(Lv.w, Lv.w, Lv.w, Lv.w);
                else if (right_block) cl::store_remote_tx(&Lrow[i - Tn * kLEntry], crank + 1, Lv, halo_bar(0, r & 1));    // the right block's left halo
            }
        }
        __syncthreads();
        if constexpr (kClustered) {                         // the one thread that reads a neighbour block's halo waits for it
            if (seg_first && !plane_first) cl::halo_wait(halo_bar(0, r & 1), halo_parity, kNumCost * 16u);
            if (right_block) cl::halo_wait(halo_bar(1, r & 1), halo_parity, kNumCost * 16u);
        }
        if constexpr (!kFull) prefetch_stale(r + 2);        // next row's handed-over state: on its way during phase B

        // ---- per cost: 7-tap sum, B = narrowT(sum / 16), min key, M = P[r+1] + B ----
        int kmin[kCols];        // integer flavours: min over (B << 4 | rank) keys, threshold folded in as a tenth key
        float fmin[kCols];      // fp32: running minimum and the rank that first reached it
        int frank[kCols];
#pragma unroll
        for (int c = 0; c < kCols; ++c) { kmin[c] = tkey; fmin[c] = 0.f; frank[c] = 0; }
        [[maybe_unused]] int held[kCols];
        T* outp = out.p;
        // Costs are visited 0, 1, .. 8 for integers (the tie order lives in the keys); for fp32 in the reference's tie
        // order 4,5,3,6,2,7,1,8,0, so that a strict '<' leaves the first of several equal minima as the winner
        // (:214-249). Visit k of fp32 has rank k. The three vectors of a cost are fetched one cost ahead of their use.
        uint4 nl = Lrow[visit_cost<kFloat>(0) - kLEntry], no = kFloat ? Lrow[visit_cost<kFloat>(0)] : as_words(own0), nr = Lrow[visit_cost<kFloat>(0) + kLEntry];
#pragma unroll
        for (int k = 0; k < kNumCost; ++k) {
            const int i = visit_cost<kFloat>(k);
            const uint4 ql = nl, qo = no, qr = nr;
            if (k + 1 < kNumCost) { const int i1 = visit_cost<kFloat>(k + 1); nl = Lrow[i1 - kLEntry]; no = Lrow[i1]; nr = Lrow[i1 + kLEntry]; }
            I Lw[kWin];             // L of columns x0-4 .. x0+7; pixel c reads indices c+1 .. c+7
            {
                I a[4], b[4], c4[4];
                from_words(ql, a); from_words(qo, b); from_words(qr, c4);
#pragma unroll
                for (int e = 0; e < 4; ++e) { Lw[e] = a[e]; Lw[4 + e] = b[e]; Lw[8 + e] = c4[e]; }
            }
            I B4[kCols];
            if constexpr (kFloat) {
#pragma unroll
                for (int c = 0; c < kCols; ++c) {
                    float s = __fadd_rn(Lw[c + 1], Lw[c + 2]);                     // ((((((m3+m2)+m1)+c)+p1)+p2)+p3) / 16  (:152)
                    s = __fadd_rn(s, Lw[c + 3]); s = __fadd_rn(s, Lw[c + 4]); s = __fadd_rn(s, Lw[c + 5]);
                    s = __fadd_rn(s, Lw[c + 6]); s = __fadd_rn(s, Lw[c + 7]);
                    B4[c] = __fmul_rn(s, 0.0625f);
                    if (k == 0 || B4[c] < fmin[c]) { fmin[c] = B4[c]; frank[c] = k; }
                }
            } else {
                int s = (Lw[1] + Lw[2] + Lw[3]) + (Lw[4] + Lw[5] + Lw[6]) + Lw[7];
#pragma unroll
                for (int c = 0; c < kCols; ++c) {
                    if (c > 0) s += Lw[c + 7] - Lw[c];
                    // narrowT(s / 16) << 4 | rank: wrapped (:152), or clamped for the SSE2 flavour (SangNom2_SSE2.cpp:807)
                    const unsigned kept = kSat ? min((unsigned)s, keymask | 15u) & keymask : (unsigned)s & keymask;
                    const int key = (int)(kept | (unsigned)rank_of(i));
                    B4[c] = key >> 4;
                    if (k & 1) kmin[c] = (int)__vimin3_u32((unsigned)kmin[c], (unsigned)held[c], (unsigned)key);
                    else if (k == kNumCost - 1) kmin[c] = min(kmin[c], key);
                    else held[c] = key;
                }
            }
#pragma unroll
            for (int c = 0; c < kCols; ++c) M[i][c] = add2(B4[c], M[i][c]);
            // hand the blurred row to the next pass of this frame
            if constexpr (kExport) {
                if (out.p != nullptr) *reinterpret_cast<V*>(outp + (size_t)i * out.stride) = pack4(B4, T());
            }
        }

        // ---- interpolate the picture row between K[r-1] and K[r] ----
        if (pixels && (kPair || r == n - 1)) {
            const T* const up = slot_ptr(sm1) + kPadS + lx - 3;             // sample x0-3 of K[r-1]
            const T* const dn = slot_ptr(s0) + kPadS + lx + 3;              // sample x0+3 of K[r]
            Tap3 tu, td;
            t3_get(ph3_prev, tu);
            t3_get(ph3, td);
            I px[kCols];
            // the winning direction's two operands, fetched by index (the staged rows carry their replicated edges)
#pragma unroll
            for (int c = 0; c < kCols; ++c) {
                int rank;
                if constexpr (kFloat) rank = fmin[c] > thr_f ? 0 : frank[c];
                else rank = kmin[c] & 15;
                const int d3 = tap_index(rank);                             // d + 3
                I a = (I)up[c + d3], b = (I)dn[c - d3];
                if (rank == 1) { a = tu.b[c]; b = td.f[c]; }
                if (rank == 2) { a = tu.f[c]; b = td.b[c]; }
                px[c] = mean2(a, b);
            }
            const int y = t.offset + 2 * (r - 1);
            store4(y + 1, pack4(px, T()));
            if (t.copy_kept) store4(y, own4(sm1));
            if (!kPair) {                                                   // K[r] is the last kept row
                if (t.offset == 0 && !t.no_border) store4(t.height - 1, own4(s0));
                if (t.copy_kept) store4(y + 2, own4(s0));
            }
        }
    };
    // The rows my warp exports are two ranges known up front (region B: rows b_r0..b_r1, all columns; region A: rows
    // 1..a_rows for the warps that reach past a_x0).
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
    auto export_row = [&](int r, StateRow<T>& out) -> bool {
        out = StateRow<T>{ nullptr, 0 };
        if (!((r >= ex_b0 && r <= ex_b1) || r <= ex_a1)) return false;
        if (x0 < t.export_cone - 3 * r) out = state_row<T>(t.out, r, x0, S);      // beyond it nothing downstream reads the row
        return true;
    };
    auto sweep = [&](auto full) {
        int r = 1, ph3 = 1, s1 = 2;
        unsigned par1 = 0u;
        StateRow<T> out;
        auto advance = [&] {
            ph3 = ph3 == 2 ? 0 : ph3 + 1;
            if (++s1 == kRing) { s1 = 0; par1 ^= 1u; }
        };
        for (; r <= n - 2 && r <= r_last; ++r) {                            // rows whose lower neighbour row is a pair row
            if (export_row(r, out)) row_step(full, std::true_type{}, std::true_type{}, r, ph3, s1, par1, out);
            else row_step(full, std::true_type{}, std::false_type{}, r, ph3, s1, par1, out);
            advance();
        }
        for (; r <= r_last; ++r) {                                          // the last picture row and rows swept for the next pass only
            if (export_row(r, out)) row_step(full, std::false_type{}, std::true_type{}, r, ph3, s1, par1, out);
            else row_step(full, std::false_type{}, std::false_type{}, r, ph3, s1, par1, out);
            advance();
        }
    };
    // warp-uniform choice, so that a warp never splits over the copies of the row barrier
    // warp-uniform choice, so that a warp never splits over the copies of the row barrier
    if (warp_full) sweep(std::true_type{}); else sweep(std::false_type{});
#ifdef SN_HOST_EMULATION
    if (r_last < R) {                                                       // left before the last row: out of the barriers
        emul::bar->arrive_and_drop();
        if (kClustered) emul::cluster_bar->arrive_and_drop();
    }
#endif
}

}  // namespace wide
}  // namespace sn
