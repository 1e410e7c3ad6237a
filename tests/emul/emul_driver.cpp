// Runs the device kernels of csrc/*.cuh on the CPU through tests/emul/cuda_emul.h, one emulated
// thread block per plane pass, with the same per-frame pass planning the C-ABI layer uses
// (csrc/sangnom_plan.h). TEST INFRASTRUCTURE ONLY - see cuda_emul.h.
#include "cuda_emul.h"

#include "sangnom_plan.h"
#include "host_copy_pool.h"
#include "sangnom_u8.cuh"
#include "sangnom_wide.cuh"
#include "sangnom_turn.cuh"

#include <cstdio>

extern "C" {

// One frame: up to 3 processed planes, in place (kept field already in the dst planes).
// planes[i]: pointer to row 0; pitch in BYTES; returns 0, or -1 for unsupported geometry.
int emul_frame(int sample_bytes, int nplanes, void* const* planes, const long long* pitch_bytes, const void* const* srcs,
               const long long* src_pitch_bytes, const int* widths,
               const int* heights, const int* offsets, const float* thresholds, int pool_width, int pool_height, int cluster,
               void* carry_in, void* carry_out, int saturate)
{
    // carry_in/carry_out != NULL: persistent-pool mode, the pool state (plan_carry_bytes) before and after this frame
    const int S = (pool_width + 31) & ~31, Hb = (pool_height + 1) >> 1;
    sn::PassGeometry geo[3];
    for (int q = 0; q < nplanes; ++q) { geo[q] = sn::PassGeometry{}; geo[q].width = widths[q]; geo[q].kept_rows = heights[q] / 2; }
    const bool persistent = carry_in != nullptr && carry_out != nullptr;
    const size_t state_bytes = sn::plan_frame_passes(geo, nplanes, S, Hb, sample_bytes, persistent);
    std::vector<char> state(state_bytes + 256, (char)0x5A);      // poisoned: a read of a never-written cell is visible
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(state.data()) + 255) & ~(uintptr_t)255);
    for (int q = 0; q < nplanes; ++q) {
        sn::plan_place_state(geo[q].in, base);
        sn::plan_place_state(geo[q].out, base);
    }
    if (persistent && nplanes > 0) sn::plan_attach_carry(geo[0].in, geo[nplanes - 1].out, carry_in, carry_out, Hb);
    for (int q = 0; q < nplanes; ++q) {
        sn::PlaneTask t{};
        t.plane = planes[q];
        t.pitch = pitch_bytes[q] / sample_bytes;
        if (srcs && srcs[q]) {                  // out of place: kept rows come from a packed field buffer
            t.src = srcs[q];
            t.src_pitch = src_pitch_bytes[q] / sample_bytes;
            t.copy_kept = 1;
        } else {                                // in place: the kept field is already in the dst plane
            t.src = static_cast<char*>(planes[q]) + (long long)offsets[q] * pitch_bytes[q];
            t.src_pitch = 2 * t.pitch;
            t.copy_kept = 0;
        }
        t.width = widths[q]; t.height = heights[q]; t.offset = offsets[q];
        t.kept_rows = geo[q].kept_rows; t.sweep_rows = geo[q].sweep_rows; t.cone = geo[q].cone; t.export_cone = geo[q].export_cone;
        t.thr_f = thresholds[q];
        t.thr_i = sample_bytes == 1 ? ((int)thresholds[q] & 0xFF) : ((int)thresholds[q] & 0xFFFF);
        t.in = geo[q].in; t.out = geo[q].out;
        const bool narrow = widths[q] + 8 <= S;
        sn::LaunchGeometry g = sn::make_geometry(S, Hb, saturate != 0, narrow);
        if (sample_bytes == 2) g.key_mask = 0xFFFF0u;                  // like launch_wide()
        const unsigned G = (unsigned)cluster;
        const int cols = sample_bytes == 1 ? sn::u8k::kCols : sn::wide::kCols;
        if (S % (int)(G * cols) != 0) return -1;
        const int seg = S / (int)G;
        unsigned threads = (unsigned)(seg / cols);
        const bool spare_threads = G == 1 && narrow && threads < 256;
        if (spare_threads) threads = std::max(threads, std::min(256u, ((threads + 31u) & ~31u) + 32u));      // like the launcher
        // one block per plane runs the unclustered instantiation, like the launcher (it alone has the spare-thread map)
        auto run = [&](size_t smem, auto kernel) { emul::run_cluster(0, G, threads, smem, [&] { kernel(&t, g, seg); }); };
        auto pick = [&](auto clustered, auto sat, auto spare) {
            constexpr bool kC = decltype(clustered)::value, kS = decltype(sat)::value, kP = decltype(spare)::value;
            if (sample_bytes == 1) run(sn::u8k::smem_bytes(seg), sn::u8k::sangnom_u8_row_sweep<1024, 1, kC, kS, kP>);
            else if (sample_bytes == 2) run(sn::wide::smem_bytes<uint16_t>(seg), sn::wide::sangnom_wide_row_sweep<uint16_t, 1024, 1, kC, kS, kP>);
            else run(sn::wide::smem_bytes<float>(seg), sn::wide::sangnom_wide_row_sweep<float, 1024, 1, kC, false, kP>);
        };
        using Y = std::true_type;
        using N = std::false_type;
        if (G == 1) {
            if (spare_threads) { if (saturate) pick(N{}, Y{}, Y{}); else pick(N{}, N{}, Y{}); }
            else { if (saturate) pick(N{}, Y{}, N{}); else pick(N{}, N{}, N{}); }
        } else {
            if (saturate) pick(Y{}, Y{}, N{}); else pick(Y{}, N{}, N{});
        }
    }
    return 0;
}

// One plane through the turn kernel (sangnom_turn.cuh): kind 0 transpose, 1 clockwise, 2 counter-clockwise.
int emul_turn(int sample_bytes, int kind, const void* src, long long src_pitch, void* dst, long long dst_pitch, int width, int height)
{
    const int TS = sn::turn::tile_side(sample_bytes);
    sn::turn::TurnBatch batch{};
    sn::turn::TurnTask& t = batch.t[0];
    t.src = src; t.dst = dst; t.src_pitch = src_pitch; t.dst_pitch = dst_pitch; t.width = width; t.height = height;
    t.tiles_x = (width + TS - 1) / TS;
    t.first_block = 0;
    const int tiles = t.tiles_x * ((height + TS - 1) / TS);
    const int blocks = (tiles + sn::turn::kTilesPerBlock - 1) / sn::turn::kTilesPerBlock;
    const int fr = kind == 2, fc = kind == 1;
    for (int b = 0; b < blocks; ++b) {
        auto body = [&] {
            if (sample_bytes == 1) sn::turn::sangnom_turn_planes<1>(batch, 1, tiles, fr, fc);
            else if (sample_bytes == 2) sn::turn::sangnom_turn_planes<2>(batch, 1, tiles, fr, fc);
            else sn::turn::sangnom_turn_planes<4>(batch, 1, tiles, fr, fc);
        };
        emul::run_block((unsigned)b, sn::turn::kThreads, sn::turn::smem_bytes(sample_bytes), body);
    }
    return 0;
}

// The host copy pool of libsangnom_cuda on its own: `rounds` batches of random strided row copies (sizes from a few
// bytes to several MB per batch, so that both the single-threaded short cut and the worker path run) against plain
// memcpy. Returns 0, or the number of the first batch that differs.
int emul_copy_pool_selftest(int threads, int rounds, unsigned seed)
{
    sn_host::CopyPool pool(threads > 0 ? threads - 1 : 0);
    uint64_t state = seed * 2654435761u + 1;
    auto rnd = [&](uint32_t n) { state = state * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)((state >> 33) % n); };
    for (int round = 1; round <= rounds; ++round) {
        const int njobs = 1 + (int)rnd(24);
        std::vector<std::vector<char>> src((size_t)njobs), dst((size_t)njobs), ref((size_t)njobs);
        std::vector<sn_host::RowCopy> jobs;
        for (int j = 0; j < njobs; ++j) {
            const size_t row = 1 + rnd(round % 3 == 0 ? 9000 : 300);
            const int rows = 1 + (int)rnd(round % 3 == 0 ? 700 : 40);
            const ptrdiff_t sp = (ptrdiff_t)(row + rnd(64)), dp = (ptrdiff_t)(row + rnd(64));
            src[(size_t)j].resize((size_t)sp * rows);
            dst[(size_t)j].assign((size_t)dp * rows, (char)0x5A);
            ref[(size_t)j].assign((size_t)dp * rows, (char)0x5A);
            for (char& c : src[(size_t)j]) c = (char)rnd(256);
            for (int y = 0; y < rows; ++y) std::memcpy(ref[(size_t)j].data() + (size_t)y * dp, src[(size_t)j].data() + (size_t)y * sp, row);
            jobs.push_back(sn_host::RowCopy{ dst[(size_t)j].data(), src[(size_t)j].data(), dp, sp, row, rows });
        }
        pool.run(jobs);
        for (int j = 0; j < njobs; ++j)
            if (dst[(size_t)j] != ref[(size_t)j]) return round;
    }
    return 0;
}

// The host entry's chunk dealing (sangnom_plan.h plan_chunks): writes (first, last, pipeline) triples, returns the count.
int emul_deal_chunks(long long nframes, long long chunk_frames, int npipelines, long long* next, long long* out, int max_out)
{
    size_t nx = (size_t)*next;
    int n = 0;
    sn::plan_chunks((size_t)nframes, (size_t)chunk_frames, npipelines, nx, [&](const sn::ChunkDeal& d) {
        if (n < max_out) { out[3 * n] = (long long)d.first; out[3 * n + 1] = (long long)d.last; out[3 * n + 2] = d.pipeline; }
        ++n;
    });
    *next = (long long)nx;
    return n;
}

size_t emul_carry_bytes(int sample_bytes, int pool_width, int pool_height)
{
    return sn::plan_carry_bytes((pool_width + 31) & ~31, (pool_height + 1) >> 1, sample_bytes);
}

}  // extern "C"
