// AviSynth plugin shim: SangNom2(...) and legacy SangNom(...) on libsangnom_cuda.
//
// Mirrors the script surface of /root/reference/src/SangNom2.cpp:
//   AvisynthPluginInit3 + the two AddFunction signatures            :474-484
//   Create_SangNom2 / Create_SangNom: defaults, checks, messages    :399-472
//   ctor: threshold scaling, dh doubles vi.height                   :275-288
//   GetFrame: field offset by order/parity, plane loop              :332-397
// The per-plane compute (prepare / blur / finalize, :74-273) is NOT here: it runs on the GPU
// behind sangnom_cuda_process_planes. There is no CPU fallback; if the device library cannot
// create a context the filter constructor raises a script error.
#include "sangnom2_filter.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

int env_int(const char* name, int def, int lo, int hi)
{
    const char* v = std::getenv(name);
    if (!v || !*v) return def;
    const int x = std::atoi(v);
    return x < lo ? lo : (x > hi ? hi : x);
}

const int kPlaneIds[3] = { PLANAR_Y, PLANAR_U, PLANAR_V };

}  // namespace

SangNom2::SangNom2(PClip _child, int order, int aa, int aac, int threads, bool dh, bool luma, bool chroma, int opt, IScriptEnvironment* env)
    : GenericVideoFilter(_child), order_(order), dh_(dh), aaf_{ 0.0f, 0.0f, 0.0f }, process_plane_{ luma, chroma, chroma }
{
    (void)threads;   // dummy parameter, as in the reference (README.md:40-41)
    // opt is validated by the factory. There is one code path on the device; opt picks the ARITHMETIC flavour by the
    // reference's own dispatch rule (:312): opt=1, or opt<0 on a host that reports SSE2, is the SSE2 path's arithmetic
    // (narrowing saturates, SangNom2_SSE2.cpp:449-517,761,807); opt=0, or opt<0 without SSE2, is the C++ path's
    // (narrowing wraps, :63-64,:152). So a script with default arguments gives what the stock reference gives.
    has_at_least_v8_ = env->FunctionExists("propShow");
    sample_bytes_ = vi.ComponentSize();
    plane_count_ = vi.NumComponents() < 3 ? vi.NumComponents() : 3;
    has_alpha_ = vi.NumComponents() == 4;

    const int strength[3] = { aa, aac, aac };
    for (int i = 0; i < plane_count_; ++i)
        aaf_[i] = sangnom_cuda_threshold(strength[i], vi.BitsPerComponent(), sample_bytes_);

    if (dh_) vi.height *= 2;

    sn_config cfg{};
    cfg.abi_version = SANGNOM_CUDA_ABI_VERSION;
    cfg.device = env_int("SANGNOM_B200_DEVICE", 0, 0, 1023);
    // SANGNOM_B200_DEVICES=all | 0,2,3: one pipeline per listed GPU behind this one filter instance, chunks of frames
    // dealt round-robin (what the reference gets from MT_MULTI_INSTANCE clones, SangNom2.h:63-66)
    if (const char* v = std::getenv("SANGNOM_B200_DEVICES")) {
        if (std::strcmp(v, "all") == 0) cfg.device = SN_DEVICE_ALL;
        else
            for (const char* p = v; *p;) {
                char* end = nullptr;
                const long d = std::strtol(p, &end, 10);
                if (end == p) break;
                if (d >= 0 && d < 64) cfg.device_mask |= 1ull << d;
                p = *end == ',' ? end + 1 : end;
            }
    }
    cfg.sample_type = sample_bytes_;
    cfg.pool_width = vi.width;       // pool geometry comes from the OUTPUT luma size (:287-288)
    cfg.pool_height = vi.height;
    // frames fetched and processed per cache miss on sequential access: enough to keep the device pipeline busy, but
    // bounded by the memory the finished frames occupy until they are served (about 1 GB per batch by default)
    const long long frame_bytes = (long long)vi.width * vi.height * sample_bytes_ * (plane_count_ == 1 ? 2 : 3) / 2 + 1;
    const int default_batch = (int)std::max(4LL, std::min(128LL, (1024LL << 20) / frame_bytes));
    batch_frames_ = env_int("SANGNOM_B200_BATCH", default_batch, 1, 256);
    // A batch is cut into four chunks (the library keeps four chunks in flight per GPU: host copies | H2D | kernels |
    // D2H). Smaller chunks lose: a chunk's kernels take about as long for 4 frames as for 40 (the row sweep is a chain
    // of H/2 dependent steps, frames only add blocks beside it), so fewer frames per chunk is fewer frames per unit of
    // kernel latency - measured 13.4k / 11.9k / 10.0k / 6.7k frames/s at 1080p for 4 / 8 / 16 / 32 chunks per batch;
    // larger ones (2 or 1 chunk per batch) stop overlapping the stages: 11.2k / 8.1k. SANGNOM_B200_INFLIGHT overrides.
    cfg.max_frames_in_flight = env_int("SANGNOM_B200_INFLIGHT", 4 * ((batch_frames_ + 3) / 4), 1, 4096);
    // SANGNOM_B200_PERSISTENT=1: keep the scratch-pool state from frame to frame like one long-lived reference
    // instance (bit-compatible with a sequential single-instance reference run where that is not frame-pure:
    // widths that are not a multiple of 32, luma=false with subsampled chroma). Frames then run one after another.
    if (env_int("SANGNOM_B200_PERSISTENT", 0, 0, 1)) cfg.flags |= SN_FLAG_PERSISTENT_POOL;
    if (opt == 1 || (opt < 0 && (env->GetCPUFlags() & CPUF_SSE2))) cfg.flags |= SN_FLAG_SATURATE;
    // SANGNOM_B200_PREFETCH: batches submitted ahead on sequential reads. 0: none (every child frame is requested only
    // when it is needed); 1: the next batch runs while the host consumes the finished one; 2 (default): one more
    // behind it, so that its upload overlaps the kernels and the download of the batch before - a consumer faster than
    // a batch's latency (a benchmark, a fast encoder) otherwise sees the device pipeline drain after every batch.
    prefetch_depth_ = env_int("SANGNOM_B200_PREFETCH", 2, 0, 4);
    // SANGNOM_B200_PIN_MB: budget for pinning the host's recycled frame buffers (0 = never). Unset: 2 GB, and only frame
    // buffers of 16 MB and more are pinned - pinning costs ~4 ms per buffer whatever its size, once, which a few dozen
    // large buffers repay (+5 % frames/s at 2160p fp32) and a few hundred small ones do not (+2 % at 1080p 8-bit
    // after seconds of warm-up; profiles/README.md).
    pin_min_bytes_ = std::getenv("SANGNOM_B200_PIN_MB") ? 0 : (size_t)16 << 20;
    pin_budget_ = (size_t)env_int("SANGNOM_B200_PIN_MB", 2048, 0, 1 << 20) << 20;
    if (sangnom_cuda_create(&cfg, &ctx_) != SN_OK)
        env->ThrowError("SangNom2: %s", sangnom_cuda_last_error(nullptr));
}

SangNom2::~SangNom2()
{
    for (const Pending& p : pending_) sangnom_cuda_wait(ctx_, p.ticket);   // nothing may still write into the frames
    for (auto& kv : pins_)
        if (kv.second.pinned) sangnom_cuda_host_unpin(ctx_, const_cast<void*>(kv.first));
    pins_.clear();
    pending_.clear();
    ready_.clear();
    sangnom_cuda_destroy(ctx_);
}

void SangNom2::note_frame_buffer(const PVideoFrame& f)
{
    if (pin_budget_ == 0) return;
    VideoFrameBuffer* const vfb = f->GetFrameBuffer();
    const void* const base = vfb->GetReadPtr();
    const size_t bytes = (size_t)vfb->GetDataSize();
    if (!base || bytes == 0 || bytes > pin_budget_ || bytes < pin_min_bytes_) return;
    PinEntry& e = pins_[base];
    if (e.bytes != bytes) {                                 // a different buffer at an address seen before
        if (e.pinned) { sangnom_cuda_host_unpin(ctx_, const_cast<void*>(base)); pinned_bytes_ -= e.bytes; }
        e = PinEntry{};
        e.bytes = bytes;
    }
    e.last_use = ++pin_clock_;
    if (e.pinned || ++e.seen < 2) return;                   // pinned already, or seen for the first time: may be a one-off
    while (pinned_bytes_ + bytes > pin_budget_) {           // make room: unpin the least recently used buffer ...
        auto lru = pins_.end();
        for (auto it = pins_.begin(); it != pins_.end(); ++it)
            if (it->second.pinned && (lru == pins_.end() || it->second.last_use < lru->second.last_use)) lru = it;
        // ... but only one that has dropped out of circulation (not touched for several batches). If every pinned
        // buffer is still in use the working set simply does not fit the budget: this buffer stays pageable, rather
        // than pinning and unpinning in turns (unpinning costs more than the staging copy it saves).
        if (lru == pins_.end() || lru->second.last_use + 8ull * 2 * (uint64_t)batch_frames_ > pin_clock_) { --e.seen; return; }
        sangnom_cuda_host_unpin(ctx_, const_cast<void*>(lru->first));
        pinned_bytes_ -= lru->second.bytes;
        pins_.erase(lru);
    }
    if (sangnom_cuda_host_pin(ctx_, const_cast<void*>(base), bytes) == SN_OK) { e.pinned = true; pinned_bytes_ += bytes; }
    else e.seen = -1000;                                    // the driver refused: leave this one pageable
    if (pins_.size() > 4096)                                // forget stale one-off entries
        for (auto it = pins_.begin(); it != pins_.end();) it = (!it->second.pinned && it->second.last_use + 2048 < pin_clock_) ? pins_.erase(it) : std::next(it);
}

// 0 keeps the top field (rows 0,2,..), 1 the bottom field (reference :336-341).
int SangNom2::field_offset(int n)
{
    switch (order_) {
        case 0: return child->GetParity(n) ? 0 : 1;
        case 1: return 0;
        default: return 1;
    }
}

void SangNom2::start_batch(int first, int count, IScriptEnvironment* env)
{
    Pending batch;
    std::vector<PVideoFrame>& srcs = batch.srcs;
    std::vector<PVideoFrame>& dsts = batch.dsts;
    srcs.assign((size_t)count, PVideoFrame());
    dsts.assign((size_t)count, PVideoFrame());
    std::vector<sn_plane_job> jobs;
    jobs.reserve((size_t)count * 5);
    for (int k = 0; k < count; ++k) {
        const int n = first + k;
        const int offset = field_offset(n);
        srcs[k] = child->GetFrame(n, env);
        dsts[k] = has_at_least_v8_ ? env->NewVideoFrameP(vi, &srcs[k]) : env->NewVideoFrame(vi);
        note_frame_buffer(srcs[k]);
        note_frame_buffer(dsts[k]);
        for (int i = 0; i < plane_count_; ++i) {
            const int plane = kPlaneIds[i];
            sn_plane_job jb{};
            jb.src = srcs[k]->GetReadPtr(plane);
            jb.src_pitch = srcs[k]->GetPitch(plane);
            jb.dst = dsts[k]->GetWritePtr(plane);
            jb.dst_pitch = dsts[k]->GetPitch(plane);
            jb.width = srcs[k]->GetRowSize(plane) / sample_bytes_;
            jb.dst_height = dsts[k]->GetHeight(plane);
            jb.offset = offset;
            // dh forces every plane to be processed (:361-366); otherwise a disabled plane is copied (:369-374)
            jb.mode = dh_ ? SN_MODE_DH : (process_plane_[i] ? SN_MODE_FIELD : SN_MODE_COPY);
            jb.threshold = aaf_[i];
            jb.plane = i;
            jb.frame = n;
            jobs.push_back(jb);
        }
        if (has_alpha_) {
            // The reference leaves the alpha plane of the new frame unwritten (:346-348). We copy it;
            // for dh every source row is written to both output rows of its pair.
            const int reps = dh_ ? 2 : 1;
            for (int r = 0; r < reps; ++r) {
                sn_plane_job jb{};
                jb.src = srcs[k]->GetReadPtr(PLANAR_A);
                jb.src_pitch = srcs[k]->GetPitch(PLANAR_A);
                jb.dst = dsts[k]->GetWritePtr(PLANAR_A) + (ptrdiff_t)r * dsts[k]->GetPitch(PLANAR_A);
                jb.dst_pitch = (ptrdiff_t)dsts[k]->GetPitch(PLANAR_A) * reps;
                jb.width = srcs[k]->GetRowSize(PLANAR_A) / sample_bytes_;
                jb.dst_height = srcs[k]->GetHeight(PLANAR_A);
                jb.mode = SN_MODE_COPY;
                jb.plane = 3;
                jb.frame = n;
                jobs.push_back(jb);
            }
        }
    }
    // the job array is copied by the library; the frames themselves stay referenced in pending_
    if (sangnom_cuda_submit(ctx_, jobs.data(), (int)jobs.size(), &batch.ticket) != SN_OK)
        env->ThrowError("SangNom2: %s", sangnom_cuda_last_error(ctx_));
    batch.first = first;
    batch.count = count;
    pending_.push_back(std::move(batch));
}

// Wait for the oldest batch in flight and file its frames. A failed batch fails this call only if it holds the frame
// being asked for (`wanted`); otherwise its frames are simply not there and are computed again when somebody asks.
void SangNom2::finish_oldest_batch(int wanted, IScriptEnvironment* env)
{
    if (pending_.empty()) return;
    Pending p = std::move(pending_.front());
    pending_.pop_front();
    const int rc = sangnom_cuda_wait(ctx_, p.ticket);
    if (rc == SN_OK)
        for (int k = 0; k < p.count; ++k) ready_[p.first + k] = p.dsts[(size_t)k];
    else if (wanted >= p.first && wanted < p.first + p.count)
        env->ThrowError("SangNom2: %s", sangnom_cuda_last_error(ctx_));
}

PVideoFrame __stdcall SangNom2::GetFrame(int n, IScriptEnvironment* env)
{
    std::lock_guard<std::mutex> lk(mu_);
    const int last = vi.num_frames > 0 ? vi.num_frames - 1 : n;
    auto window = [&](int first, int want) {            // frames [first, first+count) still to compute, clip end respected
        int count = want;
        if (first + count - 1 > last) count = last - first + 1;
        for (int k = 1; k < count; ++k)
            if (ready_.count(first + k)) { count = k; break; }     // frames already finished are not recomputed
        return count < 1 ? 0 : count;
    };
    auto in_flight = [&](int f) {
        for (const Pending& p : pending_)
            if (f >= p.first && f < p.first + p.count) return true;
        return false;
    };
    auto hit = ready_.find(n);
    if (hit == ready_.end()) {
        if (in_flight(n)) {
            while (ready_.find(n) == ready_.end() && !pending_.empty()) finish_oldest_batch(n, env);   // a prefetched batch holds it
        } else {
            while (!pending_.empty()) finish_oldest_batch(n, env);     // a seek: what is in flight comes home first
        }
        hit = ready_.find(n);
        if (hit == ready_.end()) {
            // Sequential pulls (the normal frameserver pattern) are served in batches so the GPU sees
            // many planes per launch; a seek falls back to a single frame.
            const bool sequential = (n == last_request_ + 1) || (n == 0 && last_request_ == -2);
            int count = window(n, sequential ? batch_frames_ : 1);
            if (count < 1) count = 1;
            while (!pending_.empty()) finish_oldest_batch(n, env);
            start_batch(n, count, env);
            finish_oldest_batch(n, env);
            hit = ready_.find(n);
        }
    }
    PVideoFrame out = hit->second;
    const bool sequential = (n == last_request_ + 1) || (n == 0 && last_request_ == -2);
    last_request_ = n;
    // Prefetch: on a sequential read the batches after the finished run are submitted ahead (one new batch per call),
    // so that they upload and run while the host consumes (encodes, filters further) what is ready.
    if (sequential && (int)pending_.size() < prefetch_depth_) {
        int next = (pending_.empty() ? n : pending_.back().first + pending_.back().count - 1) + 1;
        while (ready_.count(next)) ++next;
        if (next <= last && next - n <= prefetch_depth_ * batch_frames_) {
            const int count = window(next, batch_frames_);
            // Frame n is finished: a failure while fetching or submitting LATER frames must not fail this call. The
            // prefetch is dropped and the error surfaces when one of those frames is actually requested.
            if (count > 0) {
                try { start_batch(next, count, env); }
                catch (...) {}
            }
        }
    }
    // keep a bounded window of finished frames around the read position
    while ((int)ready_.size() > 2 * batch_frames_) {
        auto lo = ready_.begin();
        auto hi = std::prev(ready_.end());
        if (lo->first != n && n - lo->first >= hi->first - n) ready_.erase(lo);
        else if (hi->first != n) ready_.erase(hi);
        else break;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Factories. Checks run on the INPUT clip's VideoInfo, in the reference's order, with the
// reference's messages (SangNom2.cpp:407-422 / :446-459), including the "-1..2" wording.
namespace {

void validate(const char* name, const VideoInfo& vi, int order, int aa, const int* aac, int opt, IScriptEnvironment* env)
{
    const std::string n(name);
    if (vi.IsRGB() || !vi.IsPlanar()) env->ThrowError((n + ": clip must be in Y/YUV planar format.").c_str());
    if (vi.height % 2 != 0) env->ThrowError((n + ": height must be even.").c_str());
    if (vi.Is420() && vi.height % 4) env->ThrowError((n + ": height must be mod4.").c_str());
    if (order < 0 || order > 2) env->ThrowError((n + ": order must be between 0..2.").c_str());
    if (aa < 0 || aa > 128) env->ThrowError((n + ": aa must be between 0..128.").c_str());
    if (aac && (*aac < 0 || *aac > 128)) env->ThrowError((n + ": aac must be between 0..128.").c_str());
    if (opt < -1 || opt > 1) env->ThrowError((n + ": opt must be between -1..2.").c_str());
    if (!(env->GetCPUFlags() & CPUF_SSE2) && opt == 1) env->ThrowError((n + ": opt=1 requires SSE2.").c_str());
}

AVSValue __cdecl Create_SangNom2(AVSValue args, void*, IScriptEnvironment* env)
{
    PClip clip = args[0].AsClip();
    const VideoInfo& vi = clip->GetVideoInfo();
    const int order = args[1].AsInt(1);
    const int aa = args[2].AsInt(48);
    const int aac = args[3].AsInt(0);
    const int threads = args[4].AsInt(0);
    const bool dh = args[5].AsBool(false);
    const bool luma = args[6].AsBool(true);
    const bool chroma = args[7].AsBool(true);
    const int opt = args[8].AsInt(-1);
    validate("SangNom2", vi, order, aa, &aac, opt, env);
    return new SangNom2(clip, order, aa, aac, threads, dh, luma, chroma, opt, env);
}

// Legacy SangNom(clip, order, aa, opt): order is remapped {0->2, 1->1, 2->0} (:441,463).
// The reference's body reads the 4-entry argument array with SangNom2's indices (:443-444,
// :466-470). Observable result, reproduced here: the value the script passes as `opt` lands in
// aac (default 0 when omitted, never range-checked); threads/dh/luma/chroma/opt come from
// out-of-range subscripts and so take their defaults (0, false, true, true, -1), which also means
// the opt checks can never fire for this function and that it always runs the host's default flavour (opt=-1).
AVSValue __cdecl Create_SangNom(AVSValue args, void*, IScriptEnvironment* env)
{
    PClip clip = args[0].AsClip();
    const VideoInfo& vi = clip->GetVideoInfo();
    const int order = args[1].AsInt(1);
    const int aa = args[2].AsInt(48);
    const int aac_from_opt_slot = args[3].AsInt(0);   // the script's `opt`
    const int opt = -1;
    validate("SangNom", vi, order, aa, nullptr, opt, env);
    static const int remap[3] = { 2, 1, 0 };
    return new SangNom2(clip, remap[order], aa, aac_from_opt_slot, 0, false, true, true, opt, env);
}

}  // namespace

const AVS_Linkage* AVS_linkage = nullptr;

extern "C" __attribute__((visibility("default")))
const char* __stdcall AvisynthPluginInit3(IScriptEnvironment* env, const AVS_Linkage* const vectors)
{
    AVS_linkage = vectors;
    env->AddFunction("SangNom2", "c[order]i[aa]i[aac]i[threads]i[dh]b[luma]b[chroma]b[opt]i", Create_SangNom2, 0);
    env->AddFunction("SangNom", "c[order]i[aa]i[opt]i", Create_SangNom, 0);
    return "SangNom2";
}
