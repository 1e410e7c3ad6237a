// Anti-aliasing chain on the device (include/sangnom_cuda.h, "sangnom_cuda_chain_*"):
//   SangNom2(dh=true)  ->  turn  ->  SangNom2(dh=true)  ->  turn back
// The dominant real-world caller of the reference runs exactly this as four script filters
// (/root/reference/README.md:43-46 documents dh; the turns are AviSynth's TurnLeft/TurnRight or VapourSynth's
// Transpose), i.e. with the frame crossing host memory between every stage. Here a frame is uploaded once
// (W x H), both interpolation passes and both turns run on the GPU, and the 2W x 2H result is downloaded once.
// Each pass is the ordinary device entry of an ordinary context (same kernels, same parity contract): pass 1 on
// a context with pool (W, 2H), pass 2 on one with pool (2H, 2W).
#include "sangnom_cuda.h"
#include "sangnom_kernels.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

namespace {

thread_local std::string g_chain_create_error;

struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
    bool pinned = false;
    cudaError_t ensure(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        release();
        need = (need + 0xFFFFF) & ~(size_t)0xFFFFF;
        cudaError_t e = pinned ? cudaHostAlloc(&p, need, cudaHostAllocDefault) : cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need; else p = nullptr;
        return e;
    }
    void release() { if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); } p = nullptr; bytes = 0; }
};

constexpr int kChainSlots = 3;

struct ChainSlot {
    Buf src, mid1, mid2, mid3, out;         // W x H | W x 2H | 2H x W | 2H x 2W | 2W x 2H, all planes of the chunk
    cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
    bool busy = false;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct sn_chain {
    sn_chain_config cfg{};
    int sb = 1;
    sn_ctx* pass1 = nullptr;
    sn_ctx* pass2 = nullptr;
    cudaStream_t h2d = nullptr, d2h = nullptr, compute = nullptr;
    ChainSlot slots[kChainSlots];
    int frames_per_chunk = 1;
    sn_chain_stats stats{};
    std::string error;
    std::mutex mu;

    int fail(int code, const char* fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        error = buf;
        return code;
    }
};

#define CH_CUDA(ch, call)                                                                       \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) return (ch)->fail(SN_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

extern "C" {

const char* sangnom_cuda_chain_last_error(sn_chain* ch) { return ch ? ch->error.c_str() : g_chain_create_error.c_str(); }

void sangnom_cuda_chain_destroy(sn_chain* ch)
{
    if (!ch) return;
    cudaSetDevice(ch->cfg.device);
    cudaDeviceSynchronize();
    for (ChainSlot& s : ch->slots) {
        s.src.release(); s.mid1.release(); s.mid2.release(); s.mid3.release(); s.out.release();
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.compute_done) cudaEventDestroy(s.compute_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
    }
    if (ch->h2d) cudaStreamDestroy(ch->h2d);
    if (ch->d2h) cudaStreamDestroy(ch->d2h);
    if (ch->compute) cudaStreamDestroy(ch->compute);
    sangnom_cuda_destroy(ch->pass1);
    sangnom_cuda_destroy(ch->pass2);
    cudaGetLastError();
    delete ch;
}

int sangnom_cuda_chain_create(const sn_chain_config* cfg, sn_chain** out)
{
    if (!cfg || !out) { g_chain_create_error = "null argument"; return SN_ERR_INVALID; }
    *out = nullptr;
    if (cfg->abi_version != SANGNOM_CUDA_ABI_VERSION) { g_chain_create_error = "ABI version mismatch"; return SN_ERR_INVALID; }
    if (cfg->width <= 0 || cfg->height <= 0) { g_chain_create_error = "frame dimensions must be positive"; return SN_ERR_INVALID; }
    if (cfg->turn < SN_TURN_TRANSPOSE || cfg->turn > SN_TURN_LEFT_RIGHT) { g_chain_create_error = "turn must be 0, 1 or 2"; return SN_ERR_INVALID; }
    sn_chain* ch = new sn_chain();
    ch->cfg = *cfg;
    ch->sb = cfg->sample_type;
    sn_config c1{};
    c1.abi_version = SANGNOM_CUDA_ABI_VERSION; c1.device = cfg->device; c1.sample_type = cfg->sample_type;
    c1.pool_width = cfg->width; c1.pool_height = 2 * cfg->height;            // pass 1: W x H -> W x 2H
    sn_config c2 = c1;
    c2.pool_width = 2 * cfg->height; c2.pool_height = 2 * cfg->width;        // pass 2: 2H x W -> 2H x 2W
    int rc = sangnom_cuda_create(&c1, &ch->pass1);
    if (rc == SN_OK) rc = sangnom_cuda_create(&c2, &ch->pass2);
    if (rc != SN_OK) {
        g_chain_create_error = sangnom_cuda_last_error(nullptr);
        sangnom_cuda_chain_destroy(ch);
        return rc;
    }
    auto bail = [&](cudaError_t e, const char* what) {
        g_chain_create_error = std::string(what) + ": " + cudaGetErrorString(e);
        sangnom_cuda_chain_destroy(ch);
        return SN_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&ch->h2d, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ch->d2h, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ch->compute, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (ChainSlot& s : ch->slots) {
        if ((e = cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
        if ((e = cudaEventCreateWithFlags(&s.compute_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
        if ((e = cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    }
    // frames per chunk: a chunk has to fill the GPU with row-sweep blocks on its own (the passes of a chunk are
    // latency-bound: H or W sequential rows), i.e. about 50 frames of 3 planes; the default keeps at most ~36 GB of
    // device memory in the three slots (a frame occupies 1 + 2 + 2 + 4 + 4 = 13 times its input size)
    const double per_frame = 13.0 * 3.0 * (double)cfg->width * cfg->height * cfg->sample_type;
    const int total = cfg->max_frames_in_flight > 0 ? cfg->max_frames_in_flight : (int)std::max(3.0, std::min(150.0, 36.0e9 / per_frame));
    ch->frames_per_chunk = std::max(1, total / kChainSlots);
    *out = ch;
    return SN_OK;
}

int sangnom_cuda_chain_get_stats(sn_chain* ch, sn_chain_stats* out)
{
    if (!ch || !out) return SN_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ch->mu);
    *out = ch->stats;
    return SN_OK;
}

static int chain_process_impl(sn_chain* ch, const sn_chain_job* jobs, int njobs)
{
    if (!ch) return SN_ERR_INVALID;
    if (njobs < 0 || (njobs > 0 && !jobs)) return ch->fail(SN_ERR_INVALID, "bad job list");
    if (njobs == 0) return SN_OK;
    std::lock_guard<std::mutex> lk(ch->mu);
    CH_CUDA(ch, cudaSetDevice(ch->cfg.device));
    const int sb = ch->sb;

    // group into frames (order of first appearance), validate
    std::vector<std::vector<const sn_chain_job*>> frames;
    {
        std::map<int, size_t> index;
        for (int j = 0; j < njobs; ++j) {
            const sn_chain_job& jb = jobs[j];
            if (!jb.src || !jb.dst) return ch->fail(SN_ERR_INVALID, "job %d: null pointer", j);
            if (jb.width <= 0 || jb.height <= 0 || jb.width > ch->cfg.width || jb.height > ch->cfg.height)
                return ch->fail(SN_ERR_INVALID, "job %d: plane %dx%d does not fit the chain's %dx%d frame", j, jb.width, jb.height, ch->cfg.width, ch->cfg.height);
            if (jb.plane < 0 || jb.plane > 2) return ch->fail(SN_ERR_INVALID, "job %d: plane index %d", j, jb.plane);
            if ((jb.offset1 | jb.offset2) & ~1) return ch->fail(SN_ERR_INVALID, "job %d: offsets must be 0 or 1", j);
            if (jb.src_pitch < (ptrdiff_t)jb.width * sb || jb.dst_pitch < (ptrdiff_t)2 * jb.width * sb)
                return ch->fail(SN_ERR_INVALID, "job %d: pitch smaller than a row", j);
            auto it = index.find(jb.frame);
            if (it == index.end()) { it = index.emplace(jb.frame, frames.size()).first; frames.emplace_back(); }
            frames[it->second].push_back(&jb);
        }
    }

    const sn::TurnKind first_turn = ch->cfg.turn == SN_TURN_TRANSPOSE ? sn::kTranspose : (ch->cfg.turn == SN_TURN_RIGHT_LEFT ? sn::kTurnRight : sn::kTurnLeft);
    const sn::TurnKind second_turn = ch->cfg.turn == SN_TURN_TRANSPOSE ? sn::kTranspose : (ch->cfg.turn == SN_TURN_RIGHT_LEFT ? sn::kTurnLeft : sn::kTurnRight);

    auto drain = [&](ChainSlot& s) -> int {
        if (!s.busy) return SN_OK;
        s.busy = false;
        CH_CUDA(ch, cudaEventSynchronize(s.d2h_done));
        return SN_OK;
    };

    int status = SN_OK, slot_idx = 0;
    for (size_t next = 0; next < frames.size() && status == SN_OK;) {
        ChainSlot& s = ch->slots[slot_idx];
        slot_idx = (slot_idx + 1) % kChainSlots;
        if ((status = drain(s)) != SN_OK) break;
        const size_t first = next, last = std::min(frames.size(), next + (size_t)ch->frames_per_chunk);
        next = last;

        // device placement of every plane of the chunk in the five stage buffers (rows padded to 256 bytes)
        struct Place { const sn_chain_job* jb; size_t o_src, o1, o2, o3, o_out, p_src, p1, p2, p3, p_out; };
        std::vector<Place> pl;
        size_t b_src = 0, b1 = 0, b2 = 0, b3 = 0, b_out = 0;
        for (size_t k = first; k < last; ++k)
            for (const sn_chain_job* jb : frames[k]) {
                const size_t W = (size_t)jb->width, H = (size_t)jb->height;
                Place p{};
                p.jb = jb;
                p.p_src = align_up(W * sb, 256); p.p1 = p.p_src;              // W columns
                p.p2 = align_up(2 * H * sb, 256); p.p3 = p.p2;                // 2H columns
                p.p_out = align_up(2 * W * sb, 256);                         // 2W columns
                p.o_src = b_src; b_src += p.p_src * H;
                p.o1 = b1; b1 += p.p1 * 2 * H;
                p.o2 = b2; b2 += p.p2 * W;
                p.o3 = b3; b3 += p.p3 * 2 * W;
                p.o_out = b_out; b_out += p.p_out * 2 * H;
                pl.push_back(p);
            }
        const int np = (int)pl.size();
        cudaError_t e;
        if ((e = s.src.ensure(b_src)) != cudaSuccess || (e = s.mid1.ensure(b1)) != cudaSuccess || (e = s.mid2.ensure(b2)) != cudaSuccess ||
            (e = s.mid3.ensure(b3)) != cudaSuccess || (e = s.out.ensure(b_out)) != cudaSuccess) {
            status = ch->fail(SN_ERR_CUDA, "chain slot allocation: %s", cudaGetErrorString(e));
            break;
        }
        char* const d_src = static_cast<char*>(s.src.p);
        char* const d1 = static_cast<char*>(s.mid1.p);
        char* const d2 = static_cast<char*>(s.mid2.p);
        char* const d3 = static_cast<char*>(s.mid3.p);
        char* const d_out = static_cast<char*>(s.out.p);

        // ---- upload (once, the W x H source) ----
        for (const Place& p : pl) {
            e = cudaMemcpy2DAsync(d_src + p.o_src, p.p_src, p.jb->src, (size_t)p.jb->src_pitch, (size_t)p.jb->width * sb, (size_t)p.jb->height,
                                  cudaMemcpyHostToDevice, ch->h2d);
            if (e != cudaSuccess) { status = ch->fail(SN_ERR_CUDA, "H2D copy: %s", cudaGetErrorString(e)); break; }
            ch->stats.h2d_bytes += (uint64_t)p.jb->width * sb * p.jb->height;
        }
        if (status != SN_OK) break;
        if ((e = cudaEventRecord(s.h2d_done, ch->h2d)) != cudaSuccess || (e = cudaStreamWaitEvent(ch->compute, s.h2d_done, 0)) != cudaSuccess) {
            status = ch->fail(SN_ERR_CUDA, "event: %s", cudaGetErrorString(e)); break;
        }

        // ---- pass 1: W x H -> W x 2H ----
        std::vector<sn_plane_job> pj((size_t)np);
        for (int i = 0; i < np; ++i) {
            const Place& p = pl[i];
            sn_plane_job& j = pj[i];
            j = sn_plane_job{};
            j.src = d_src + p.o_src; j.src_pitch = (ptrdiff_t)p.p_src;
            j.dst = d1 + p.o1; j.dst_pitch = (ptrdiff_t)p.p1;
            j.width = p.jb->width; j.dst_height = 2 * p.jb->height;
            j.offset = p.jb->offset1; j.mode = SN_MODE_DH; j.threshold = p.jb->threshold; j.plane = p.jb->plane; j.frame = p.jb->frame;
        }
        if (sangnom_cuda_process_planes_device(ch->pass1, pj.data(), np, ch->compute) != SN_OK) {
            status = ch->fail(SN_ERR_CUDA, "pass 1: %s", sangnom_cuda_last_error(ch->pass1)); break;
        }
        // ---- turn: W x 2H -> 2H x W ----
        std::vector<sn::TurnPlane> tp((size_t)np);
        for (int i = 0; i < np; ++i) {
            const Place& p = pl[i];
            tp[i] = sn::TurnPlane{ d1 + p.o1, (long long)p.p1, d2 + p.o2, (long long)p.p2, p.jb->width, 2 * p.jb->height, 1 };
        }
        int nl = 0;
        if ((e = sn::launch_turn_planes(sb, tp.data(), np, first_turn, ch->compute, &nl)) != cudaSuccess) {
            status = ch->fail(SN_ERR_CUDA, "turn: %s", cudaGetErrorString(e)); break;
        }
        // ---- pass 2: 2H x W -> 2H x 2W ----
        for (int i = 0; i < np; ++i) {
            const Place& p = pl[i];
            sn_plane_job& j = pj[i];
            j.src = d2 + p.o2; j.src_pitch = (ptrdiff_t)p.p2;
            j.dst = d3 + p.o3; j.dst_pitch = (ptrdiff_t)p.p3;
            j.width = 2 * p.jb->height; j.dst_height = 2 * p.jb->width;
            j.offset = p.jb->offset2;
        }
        if (sangnom_cuda_process_planes_device(ch->pass2, pj.data(), np, ch->compute) != SN_OK) {
            status = ch->fail(SN_ERR_CUDA, "pass 2: %s", sangnom_cuda_last_error(ch->pass2)); break;
        }
        // ---- turn back: 2H x 2W -> 2W x 2H ----
        for (int i = 0; i < np; ++i) {
            const Place& p = pl[i];
            tp[i] = sn::TurnPlane{ d3 + p.o3, (long long)p.p3, d_out + p.o_out, (long long)p.p_out, 2 * p.jb->height, 2 * p.jb->width, 1 };
        }
        ch->stats.kernel_launches += nl;
        if ((e = sn::launch_turn_planes(sb, tp.data(), np, second_turn, ch->compute, &nl)) != cudaSuccess) {
            status = ch->fail(SN_ERR_CUDA, "turn: %s", cudaGetErrorString(e)); break;
        }
        ch->stats.kernel_launches += nl;           // turn launches; the passes count in their own contexts
        if ((e = cudaEventRecord(s.compute_done, ch->compute)) != cudaSuccess || (e = cudaStreamWaitEvent(ch->d2h, s.compute_done, 0)) != cudaSuccess) {
            status = ch->fail(SN_ERR_CUDA, "event: %s", cudaGetErrorString(e)); break;
        }

        // ---- download (once, the 2W x 2H result) ----
        for (const Place& p : pl) {
            e = cudaMemcpy2DAsync(p.jb->dst, (size_t)p.jb->dst_pitch, d_out + p.o_out, p.p_out, (size_t)2 * p.jb->width * sb, (size_t)2 * p.jb->height,
                                  cudaMemcpyDeviceToHost, ch->d2h);
            if (e != cudaSuccess) { status = ch->fail(SN_ERR_CUDA, "D2H copy: %s", cudaGetErrorString(e)); break; }
            ch->stats.d2h_bytes += (uint64_t)4 * p.jb->width * sb * p.jb->height;
        }
        if (status != SN_OK) break;
        if ((e = cudaEventRecord(s.d2h_done, ch->d2h)) != cudaSuccess) { status = ch->fail(SN_ERR_CUDA, "event: %s", cudaGetErrorString(e)); break; }
        s.busy = true;
        ch->stats.frames += last - first;
    }
    for (int k = 0; k < kChainSlots; ++k) {
        ChainSlot& s = ch->slots[(slot_idx + k) % kChainSlots];
        if (status == SN_OK) status = drain(s);
        else if (s.busy) { cudaEventSynchronize(s.d2h_done); s.busy = false; }
    }
    if (status != SN_OK) { cudaStreamSynchronize(ch->h2d); cudaStreamSynchronize(ch->compute); cudaStreamSynchronize(ch->d2h); cudaGetLastError(); }
    {
        sn_stats a{}, b{};
        sangnom_cuda_get_stats(ch->pass1, &a);
        sangnom_cuda_get_stats(ch->pass2, &b);
        ch->stats.pass_kernel_launches = a.kernel_launches + b.kernel_launches;
    }
    return status;
}

int sangnom_cuda_chain_process(sn_chain* ch, const sn_chain_job* jobs, int njobs)
{
    try { return chain_process_impl(ch, jobs, njobs); }
    catch (const std::bad_alloc&) { if (ch) ch->error = "out of host memory"; return SN_ERR_NOMEM; }
    catch (...) { if (ch) ch->error = "C++ exception"; return SN_ERR_INVALID; }
}

int sangnom_cuda_turn_planes_device(int sample_type, int kind, const sn_turn_plane* planes, int nplanes, void* cuda_stream)
{
    if (nplanes < 0 || (nplanes > 0 && !planes) || kind < 0 || kind > 2) return SN_ERR_INVALID;
    if (nplanes == 0) return SN_OK;
    std::vector<sn::TurnPlane> tp((size_t)nplanes);
    for (int i = 0; i < nplanes; ++i)
        tp[i] = sn::TurnPlane{ planes[i].src, (long long)planes[i].src_pitch, planes[i].dst, (long long)planes[i].dst_pitch, planes[i].width, planes[i].height, (planes[i].flags & SN_TURN_DST_PADDING_WRITABLE) ? 1 : 0 };
    const cudaError_t e = sn::launch_turn_planes(sample_type, tp.data(), nplanes, static_cast<sn::TurnKind>(kind), static_cast<cudaStream_t>(cuda_stream));
    if (e != cudaSuccess) { cudaGetLastError(); return SN_ERR_CUDA; }
    return SN_OK;
}

}  // extern "C"
