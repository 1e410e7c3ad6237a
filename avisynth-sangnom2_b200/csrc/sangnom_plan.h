// Per-frame pass planning: pure host C++ (no CUDA), shared by the C-ABI layer and the kernel
// emulation tests.
//
// Reference behaviour being reproduced: every plane's cost recursion runs over the whole scratch
// pool (rows 1..Hb-1, all S columns; /root/reference/src/SangNom2.cpp:133-136,269-270) while only
// rows 1..n-1, columns < W are freshly written for that plane (:79-81), and the pool is shared by
// the planes of a frame and never cleared (:303-310). So pass q reads, outside its own rectangle,
// what pass q-1 left there - itself the result of q-1's recursion over what q-2 left, and so on.
// Consequences planned here:
//   * a cell pass q reads at pool row r must have been swept by every earlier pass of the frame,
//     so earlier passes may have to sweep more rows than their own picture needs (sweep_rows);
//   * pass q hands pass q+1 exactly the cells q+1 will read outside its rectangle (CostState).
//   * of those cells only the ones inside the dependency cone of a picture sample matter. The recursion reaches 3
//     columns sideways per row (:144-152), so a cell at pool row r, column c can influence a sample of a plane of
//     width W with n kept rows only if c < W + 3 (n - 1 - r) + (what the later passes need through the hand-over, 6
//     columns per pass). Every pass gets that bound as `cone`: the columns it has to work on at row r are those
//     below cone - 3r; the warps right of it retire, and nothing is exported from there. Exactness is not touched:
//     the cells left out cannot reach any output.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>

#include "sangnom_kernels.h"

namespace sn {

struct PassGeometry {
    int width;        // W: picture columns of this plane
    int kept_rows;    // n = H/2
    int sweep_rows;   // out: R
    int cone;         // out: see PlaneTask::cone
    int export_cone;  // out: see PlaneTask::export_cone
    CostState in, out;   // out: regions; pointers hold (byte offset + 1) into the frame's state scratch, 0 = empty
};

inline size_t plan_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

// passes[0..m) in processing order (Y,U,V). Returns the bytes of cost-state scratch the frame needs.
// persistent: the pool outlives the frame (one long-lived reference instance pulled sequentially): the last pass
// must leave the whole pool as the reference's last processBuffers sweep does (rows 1..Hb-1, :133-136), so it - and
// through the rule below every earlier pass - sweeps all pool rows; plan_attach_carry() wires the frame-to-frame state.
inline size_t plan_frame_passes(PassGeometry* passes, int m, int S, int Hb, int sample_bytes, bool persistent = false)
{
    for (int q = m - 1; q >= 0; --q) {
        PassGeometry& p = passes[q];
        p.in = CostState{};
        p.out = CostState{};
        p.sweep_rows = std::min(p.kept_rows - 1, Hb - 1);
        if (persistent && q == m - 1) p.sweep_rows = std::max(p.sweep_rows, Hb - 1);
        if (q + 1 < m) p.sweep_rows = std::max(p.sweep_rows, std::min(Hb - 1, passes[q + 1].sweep_rows + 1));
        // Columns of blurred row r that are still needed: need(r) = max(W [r <= n-1], what pass q+1 reads of row r,
        // need(r+1) + 3). Pass q+1 reads its stale input P[r] up to 3 columns right of what it computes at rows r-1
        // and r, so the bound grows by 6 per pass. All terms are lines of slope -3 in r. A thread must still be there
        // at row r if it owns a column of the vertical sums L[r] that a needed cell reads: column < need(r) + 3.
        p.cone = p.width + 3 * p.kept_rows;
        if (q + 1 < m) p.cone = std::max(p.cone, passes[q + 1].cone + 6);
        if (persistent) p.cone = kNoCone;
        // what pass q+1 reads of row r: the stale input of the rows it computes at r-1 and r, 3 columns beyond each
        p.export_cone = (q + 1 < m && !persistent) ? passes[q + 1].cone + 3 : kNoCone;
    }
    size_t off = 0;
    for (int q = 0; q + 1 < m; ++q) {
        PassGeometry& p = passes[q];
        const PassGeometry& nx = passes[q + 1];
        CostState st{};
        // region A: rows inside the next pass's row range, columns right of its rectangle. Stored from
        // the 8-column boundary at or left of the next pass's width so that a thread's column group is
        // one aligned vector; the reader masks by its own width.
        const int a_rows = std::min(nx.kept_rows - 1, p.sweep_rows);
        if (nx.width < S && a_rows >= 1) {
            st.a_x0 = nx.width & ~7;
            st.a_rows = a_rows;
            st.a = reinterpret_cast<void*>(off + 1);
            off += plan_align((size_t)kNumCost * (a_rows + 1) * (S - st.a_x0) * sample_bytes, 256);
        }
        // region B: rows below the next pass's rectangle that its recursion still reads (row Hb and
        // beyond are never written by anyone: they read as the pool's initial zero)
        const int b0 = std::max(nx.kept_rows, 1), b1 = std::min(std::min(nx.sweep_rows + 1, Hb - 1), p.sweep_rows);
        if (b1 >= b0) {
            st.b_r0 = b0;
            st.b_r1 = b1;
            st.b = reinterpret_cast<void*>(off + 1);
            off += plan_align((size_t)kNumCost * (b1 - b0 + 1) * S * sample_bytes, 256);
        }
        p.out = st;
        passes[q + 1].in = st;
    }
    return off;
}

// Persistent pool: the state between frames is the whole pool, rows 1..Hb-1 x S columns of the nine buffers (row 0 and
// row Hb are never written by the reference and stay zero). The first pass of a frame reads it wherever it has no
// fresh raw costs of its own, the last pass writes it. Call after plan_place_state(); carry_in != carry_out.
inline size_t plan_carry_bytes(int S, int Hb, int sample_bytes) { return Hb > 1 ? (size_t)kNumCost * (Hb - 1) * S * sample_bytes : 0; }
inline void plan_attach_carry(CostState& first_in, CostState& last_out, void* carry_in, void* carry_out, int Hb)
{
    if (Hb <= 1) return;
    first_in = CostState{ nullptr, carry_in, 0, 0, 1, Hb - 1 };
    last_out = CostState{ nullptr, carry_out, 0, 0, 1, Hb - 1 };
}

// Several devices behind one context: a batch of frames is cut into chunks of consecutive frames and the chunks are
// dealt round-robin to the devices' pipelines, continuing where the previous batch stopped (`next`). A frame is never
// split - its planes share cost state - and no data moves between devices.
struct ChunkDeal { size_t first, last; int pipeline; };        // frames [first, last) of the batch
template <typename Emit>
inline void plan_chunks(size_t nframes, size_t chunk_frames, int npipelines, size_t& next, Emit&& emit)
{
    if (chunk_frames == 0) chunk_frames = 1;
    for (size_t first = 0; first < nframes; first += chunk_frames) {
        emit(ChunkDeal{ first, std::min(nframes, first + chunk_frames), (int)next });
        next = (next + 1) % (size_t)std::max(npipelines, 1);
    }
}

inline void plan_place_state(CostState& s, char* base)
{
    if (s.a) s.a = base + (reinterpret_cast<size_t>(s.a) - 1);
    if (s.b) s.b = base + (reinterpret_cast<size_t>(s.b) - 1);
}

}  // namespace sn
