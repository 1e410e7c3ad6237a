// libsangnom_cuda C ABI (include/sangnom_cuda.h): context, frame planning, the device entry, and the front end of the
// host entry (submit / wait). The per-device host pipeline is in sangnom_pipeline.cu, the kernels behind
// sangnom_kernels.h.
//
// Host-side counterpart of the reference's GetFrame plane loop (/root/reference/src/SangNom2.cpp:346-394) for a BATCH
// of frames: the kept-field copy (:361-377) becomes an upload of the kept rows, the per-plane `process` call (:393)
// becomes one thread block (or cluster) of the row-sweep kernel, and the interpolated rows are copied back.
// There is no CPU compute path in this library: without a CUDA device every call fails.
#include "sangnom_ctx.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <new>
#include <stdexcept>

using namespace sn_host;

namespace {

thread_local std::string g_create_error;

#define SN_CUDA(ctx, call)                                          \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, #call); \
    } while (0)

// Group jobs into frames and validate them; sangnom_plan.h derives for every processed plane how many
// pool rows its pass must sweep and which blurred cost cells it hands to the next pass.
int plan_frames(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, bool device_entry, std::vector<FramePlan>& frames)
{
    const int sb = ctx->sample_bytes;
    std::map<int, size_t> index;
    for (int j = 0; j < njobs; ++j) {
        const sn_plane_job& jb = jobs[j];
        if (jb.width <= 0 || jb.dst_height <= 0) return ctx->fail(SN_ERR_INVALID, "job %d: empty plane", j);
        if (jb.plane < 0 || jb.plane > 3) return ctx->fail(SN_ERR_INVALID, "job %d: plane index %d", j, jb.plane);
        if (!jb.dst || (jb.mode != SN_MODE_INPLACE && !jb.src)) return ctx->fail(SN_ERR_INVALID, "job %d: null pointer", j);
        if (jb.mode == SN_MODE_INPLACE && !device_entry) return ctx->fail(SN_ERR_INVALID, "job %d: SN_MODE_INPLACE needs the device entry", j);
        if (jb.mode < SN_MODE_COPY || jb.mode > SN_MODE_INPLACE) return ctx->fail(SN_ERR_INVALID, "job %d: mode %d", j, jb.mode);
        const ptrdiff_t row_bytes = (ptrdiff_t)jb.width * sb;
        if (jb.dst_pitch < row_bytes || (jb.mode != SN_MODE_INPLACE && jb.src_pitch < row_bytes))
            return ctx->fail(SN_ERR_INVALID, "job %d: pitch smaller than a row", j);
        if (jb.dst_pitch % sb != 0 || (jb.mode != SN_MODE_INPLACE && jb.src_pitch % sb != 0))
            return ctx->fail(SN_ERR_INVALID, "job %d: pitch not a multiple of the sample size", j);
        if ((reinterpret_cast<uintptr_t>(jb.dst) | (jb.mode != SN_MODE_INPLACE ? reinterpret_cast<uintptr_t>(jb.src) : 0)) % (uintptr_t)sb != 0)
            return ctx->fail(SN_ERR_INVALID, "job %d: pointer not aligned to the sample size", j);
        auto it = index.find(jb.frame);
        if (it == index.end()) {
            it = index.emplace(jb.frame, frames.size()).first;
            frames.emplace_back();
            frames.back().key = jb.frame;
        }
        FramePlan& f = frames[it->second];
        if (jb.mode == SN_MODE_COPY) { f.copies.push_back(&jb); continue; }
        if (jb.offset != 0 && jb.offset != 1) return ctx->fail(SN_ERR_INVALID, "job %d: offset %d", j, jb.offset);
        if (jb.dst_height % 2 != 0) return ctx->fail(SN_ERR_INVALID, "job %d: dst_height must be even", j);
        if (jb.width > ctx->S) return ctx->fail(SN_ERR_INVALID, "job %d: width %d exceeds the pool width %d", j, jb.width, ctx->S);
        if (jb.dst_height / 2 > ctx->Hb) return ctx->fail(SN_ERR_INVALID, "job %d: height %d exceeds the pool height", j, jb.dst_height);
        if (jb.plane == 3) return ctx->fail(SN_ERR_INVALID, "job %d: the alpha plane is never interpolated; use SN_MODE_COPY", j);
        Pass p{};
        p.job = &jb;
        p.W = jb.width; p.H = jb.dst_height; p.n = jb.dst_height / 2;
        for (const Pass& q : f.passes)
            if (q.job->plane == jb.plane) return ctx->fail(SN_ERR_INVALID, "frame %d: plane %d given twice", jb.frame, jb.plane);
        f.passes.push_back(p);
    }
    for (FramePlan& f : frames) {
        std::stable_sort(f.passes.begin(), f.passes.end(), [](const Pass& a, const Pass& b) { return a.job->plane < b.job->plane; });
        const int m = (int)f.passes.size();
        sn::PassGeometry geo[3];
        for (int q = 0; q < m; ++q) { geo[q] = sn::PassGeometry{}; geo[q].width = f.passes[q].W; geo[q].kept_rows = f.passes[q].n; }
        f.state_bytes = sn::plan_frame_passes(geo, m, ctx->S, ctx->Hb, sb, ctx->persistent);
        for (int q = 0; q < m; ++q) { f.passes[q].R = geo[q].sweep_rows; f.passes[q].cone = geo[q].cone; f.passes[q].export_cone = geo[q].export_cone; f.passes[q].in = geo[q].in; f.passes[q].out = geo[q].out; }
    }
    return SN_OK;
}

}  // namespace

namespace sn_host {

// Resolve the frame's cost-state regions inside its scratch; in persistent-pool mode also hook the frame into the
// chain of pool states (frames are placed in submission order, which is the order they run in).
void place_state(sn_ctx* ctx, FramePlan& f, char* base)
{
    for (Pass& p : f.passes) { sn::plan_place_state(p.in, base); sn::plan_place_state(p.out, base); }
    if (ctx->persistent && !f.passes.empty() && ctx->carry[0].p) {
        std::lock_guard<std::mutex> lk(ctx->carry_mu);
        sn::plan_attach_carry(f.passes.front().in, f.passes.back().out, ctx->carry[ctx->carry_pos].p, ctx->carry[ctx->carry_pos ^ 1].p, ctx->Hb);
        ctx->carry_pos ^= 1;
    }
}

// Group the tasks of a set of frames into launches. Default: one launch per pass index over all frames (frames are
// independent). Persistent pool: one launch per plane pass, frame after frame, because each frame reads the pool
// state the previous one wrote.
void add_task(const sn_ctx* ctx, std::vector<std::vector<sn::PlaneTask>>& launches, size_t q, const sn::PlaneTask& t)
{
    if (ctx->persistent) { launches.emplace_back(1, t); return; }
    if (launches.size() <= q) launches.resize(q + 1);
    launches[q].push_back(t);
}

// copy_kept: 1 the kernel writes the kept rows into the plane, 0 it does not (and leaves the border row alone too: the
// host path keeps both on the host), -1 decide from the pointers (device entry: in place or not).
sn::PlaneTask make_task(const sn_ctx* ctx, const Pass& p, void* plane, size_t pitch_bytes, const void* kept0, size_t kept_step_bytes, int copy_kept)
{
    sn::PlaneTask t{};
    t.plane = plane;
    t.pitch = (long long)(pitch_bytes / ctx->sample_bytes);
    t.src = kept0;
    t.src_pitch = (long long)(kept_step_bytes / ctx->sample_bytes);
    if (copy_kept < 0) {
        // in place when the kept rows already sit at rows offset, offset+2, .. of the dst plane
        t.copy_kept = !(kept0 == static_cast<char*>(plane) + (size_t)p.job->offset * pitch_bytes && kept_step_bytes == 2 * pitch_bytes);
        t.no_border = 0;
    } else {
        t.copy_kept = copy_kept;
        t.no_border = copy_kept ? 0 : 1;
    }
    t.width = p.W; t.height = p.H; t.offset = p.job->offset;
    t.kept_rows = p.n; t.sweep_rows = p.R; t.cone = p.cone; t.export_cone = p.export_cone;
    t.thr_f = p.job->threshold;
    // `const T aaf` (reference SangNom2.cpp:162,272): float -> T for integer samples. Truncate
    // toward zero, then wrap to the container width - what x86-64 does for the (undefined in C++)
    // negative case the legacy SangNom() entry can produce.
    if (ctx->sample_bytes == 1) t.thr_i = (int)p.job->threshold & 0xFF;
    else if (ctx->sample_bytes == 2) t.thr_i = (int)p.job->threshold & 0xFFFF;
    t.in = p.in; t.out = p.out;
    return t;
}

cudaError_t upload_tasks(const std::vector<std::vector<sn::PlaneTask>>& by_pass, sn::PlaneTask* host_tasks, sn::PlaneTask* dev_tasks, cudaStream_t stream)
{
    size_t total = 0;
    for (auto& v : by_pass) { std::memcpy(host_tasks + total, v.data(), v.size() * sizeof(sn::PlaneTask)); total += v.size(); }
    if (total == 0) return cudaSuccess;
    return cudaMemcpyAsync(dev_tasks, host_tasks, total * sizeof(sn::PlaneTask), cudaMemcpyHostToDevice, stream);
}

// One kernel per entry of by_pass (all first planes, then all second planes, ...), stream-ordered so that pass q+1 of
// a frame sees pass q's cost state.
cudaError_t launch_passes(sn_ctx* ctx, const std::vector<std::vector<sn::PlaneTask>>& by_pass, sn::PlaneTask* dev_tasks, cudaStream_t stream)
{
    size_t pos = 0;
    uint64_t launches = 0, planes = 0;
    for (auto& v : by_pass) {
        if (v.empty()) continue;
        bool narrow = false;                 // any plane of this launch narrower than the pool by a thread's 8 columns or more
        for (const sn::PlaneTask& t : v) narrow = narrow || t.width + 8 <= ctx->S;
        const cudaError_t e = sn::launch_plane_tasks(ctx->sample_bytes, dev_tasks + pos, (int)v.size(), sn::make_geometry(ctx->S, ctx->Hb, ctx->saturate, narrow), stream);
        if (e != cudaSuccess) return e;
        ++launches;
        planes += v.size();
        pos += v.size();
    }
    std::lock_guard<std::mutex> lk(ctx->stats_mu);
    ctx->stats.kernel_launches += launches;
    ctx->stats.planes_processed += planes;
    return cudaSuccess;
}

}  // namespace sn_host

extern "C" {

int sangnom_cuda_abi_version(void) { return SANGNOM_CUDA_ABI_VERSION; }

const char* sangnom_cuda_last_error(sn_ctx* ctx)
{
    if (!ctx) return g_create_error.c_str();
    // the string object lives in the context; the caller reads it before its next call on this context
    std::lock_guard<std::mutex> lk(ctx->err_mu);
    return ctx->error.c_str();
}

float sangnom_cuda_threshold(int aa, int bits, int sample_type)
{
    // float arithmetic in the reference's order (SangNom2.cpp:282)
    if (sample_type < 4) return (aa * 21.0f / 16.0f) * (float)(1 << (bits - 8));
    return (aa * 21.0f / 16.0f) / 256.0f;
}

int sangnom_cuda_get_limits(int device, sn_limits* out)
{
    if (!out) return SN_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    cudaDeviceProp prop{};
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e); return SN_ERR_CUDA; }
    out->max_pool_width[1] = sn::max_pool_width(1);
    out->max_pool_width[2] = sn::max_pool_width(2);
    out->max_pool_width[4] = sn::max_pool_width(4);
    out->sm_count = prop.multiProcessorCount;
    out->compute_major = prop.major;
    out->compute_minor = prop.minor;
    return SN_OK;
}

void* sangnom_cuda_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void sangnom_cuda_host_free(void* p) { if (p) cudaFreeHost(p); }

int sangnom_cuda_host_pin(sn_ctx* ctx, void* base, size_t bytes)
{
    if (!ctx || !base || bytes == 0) return SN_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (const sn_ctx::Pinned& r : ctx->pinned)
        if (r.base == base) return r.bytes >= bytes ? SN_OK : ctx->fail(SN_ERR_INVALID, "range already pinned with a smaller size");
    cudaSetDevice(ctx->devices[0]);
    const cudaError_t e = cudaHostRegister(base, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return ctx->cuda_fail(e, "cudaHostRegister"); }
    ctx->pinned.push_back(sn_ctx::Pinned{ base, bytes });
    return SN_OK;
}

int sangnom_cuda_host_unpin(sn_ctx* ctx, void* base)
{
    if (!ctx || !base) return SN_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (size_t i = 0; i < ctx->pinned.size(); ++i)
        if (ctx->pinned[i].base == base) {
            cudaSetDevice(ctx->devices[0]);
            const cudaError_t e = cudaHostUnregister(base);
            ctx->pinned.erase(ctx->pinned.begin() + (ptrdiff_t)i);
            if (e != cudaSuccess) { cudaGetLastError(); return ctx->cuda_fail(e, "cudaHostUnregister"); }
            return SN_OK;
        }
    return ctx->fail(SN_ERR_INVALID, "range was not pinned through this context");
}

int sangnom_cuda_device_count(sn_ctx* ctx) { return ctx ? (int)ctx->devices.size() : 0; }

int sangnom_cuda_create(const sn_config* cfg, sn_ctx** out)
{
    if (!cfg || !out) { g_create_error = "null argument"; return SN_ERR_INVALID; }
    *out = nullptr;
    if (cfg->abi_version != SANGNOM_CUDA_ABI_VERSION) { g_create_error = "ABI version mismatch"; return SN_ERR_INVALID; }
    if (cfg->sample_type != SN_SAMPLE_U8 && cfg->sample_type != SN_SAMPLE_U16 && cfg->sample_type != SN_SAMPLE_F32) {
        g_create_error = "sample_type must be 1, 2 or 4"; return SN_ERR_INVALID;
    }
    if (cfg->pool_width <= 0 || cfg->pool_height <= 0) { g_create_error = "pool dimensions must be positive"; return SN_ERR_INVALID; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (libsangnom_cuda has no CPU path)";
        cudaGetLastError();
        return SN_ERR_CUDA;
    }
    // which devices
    std::vector<int> devices;
    if (cfg->device_mask != 0) {
        for (int d = 0; d < 64; ++d)
            if (cfg->device_mask >> d & 1ull) {
                if (d >= count) { g_create_error = "device_mask names a device ordinal that does not exist"; return SN_ERR_INVALID; }
                devices.push_back(d);
            }
    } else if (cfg->device == SN_DEVICE_ALL) {
        for (int d = 0; d < count; ++d) devices.push_back(d);
    } else {
        if (cfg->device < 0 || cfg->device >= count) { g_create_error = "device ordinal out of range"; return SN_ERR_INVALID; }
        devices.push_back(cfg->device);
    }
    cudaDeviceProp prop{};
    for (int d : devices) {
        e = cudaGetDeviceProperties(&prop, d);
        if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return SN_ERR_CUDA; }
        if (prop.major != 10) {
            char b[160];
            snprintf(b, sizeof b, "device %d is sm_%d%d; this library carries sm_100a code only", d, prop.major, prop.minor);
            g_create_error = b;
            return SN_ERR_CUDA;
        }
    }
    const int S = (cfg->pool_width + 31) & ~31;               // reference SangNom2.cpp:287
    const int Hb = (cfg->pool_height + 1) >> 1;               // reference SangNom2.cpp:288
    if (S > sn::max_pool_width(cfg->sample_type)) {
        char b[160];
        snprintf(b, sizeof b, "pool width %d exceeds the supported maximum %d for %d-byte samples", S, sn::max_pool_width(cfg->sample_type), cfg->sample_type);
        g_create_error = b;
        return SN_ERR_UNSUPPORTED;
    }
    if (!sn::pool_width_supported(cfg->sample_type, S)) {
        char b[200];
        snprintf(b, sizeof b, "pool width %d (from clip width %d) is not supported: 8-bit pools wider than 8192 samples must be a multiple of 64", S, cfg->pool_width);
        g_create_error = b;
        return SN_ERR_UNSUPPORTED;
    }
    sn_ctx* ctx = new sn_ctx();
    ctx->cfg = *cfg;
    ctx->sample_bytes = cfg->sample_type;
    ctx->S = S; ctx->Hb = Hb;
    ctx->saturate = (cfg->flags & SN_FLAG_SATURATE) != 0;
    ctx->persistent = (cfg->flags & SN_FLAG_PERSISTENT_POOL) != 0;
    if (ctx->persistent) devices.resize(1);                    // frames are chained: one device, strictly in order
    ctx->devices = devices;
    if (cfg->max_frames_in_flight > 0) {
        ctx->frames_in_flight = cfg->max_frames_in_flight;
    } else {
        // default: one frame per SM over the kSlots chunks in flight - chunks small enough that the pipeline's ramp
        // (first upload, last download, one kernel latency) stays short, large enough that the chunks whose kernels
        // overlap fill the GPU - and at most about 512 MB of planes per chunk, so that a batch of large frames still
        // makes enough chunks to overlap the stages and to go round several devices
        const double frame_bytes = 1.5 * (double)S * (double)cfg->pool_height * (double)cfg->sample_type;      // 4:2:0-sized estimate
        const long long by_bytes = std::max<long long>(1, (long long)(512.0e6 / frame_bytes));
        const long long chunk = std::min<long long>(std::max(1, prop.multiProcessorCount / kSlots), by_bytes);
        ctx->frames_in_flight = (int)(chunk * kSlots);
    }
    auto bail = [&](cudaError_t err, const char* what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        sangnom_cuda_destroy(ctx);
        return SN_ERR_CUDA;
    };
    // host threads for row copies: the pipelines share them out
    int copy_threads = cfg->copy_threads;
    if (copy_threads <= 0) {
        const char* v = getenv("SANGNOM_B200_COPY_THREADS");
        copy_threads = v && *v ? atoi(v) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency() / 2));
    }
    copy_threads = std::max(1, std::min(copy_threads, 64));
    const int per_pipeline = std::max(1, copy_threads / (int)devices.size());
    for (int d : devices) {
        ctx->pipelines.emplace_back(new Pipeline(ctx, d, per_pipeline));
        if ((e = ctx->pipelines.back()->init()) != cudaSuccess) return bail(e, "pipeline setup");
    }
    if ((e = cudaSetDevice(devices[0])) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&ctx->own_compute, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (int i = 0; i < kTaskRing; ++i)
        if ((e = cudaEventCreateWithFlags(&ctx->dev_task_free[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if (ctx->persistent) {
        const size_t bytes = sn::plan_carry_bytes(S, Hb, cfg->sample_type);
        for (DevBuf& c : ctx->carry)
            if (bytes) {
                if ((e = c.ensure(bytes)) != cudaSuccess) return bail(e, "pool state allocation");
                if ((e = cudaMemset(c.p, 0, bytes)) != cudaSuccess) return bail(e, "cudaMemset");      // the zero-filled pool of a new instance
            }
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bail(e, "cudaDeviceSynchronize");
    }
    for (auto& p : ctx->pipelines) p->start();
    *out = ctx;
    return SN_OK;
}

void sangnom_cuda_destroy(sn_ctx* ctx)
{
    if (!ctx) return;
    for (auto& p : ctx->pipelines) p->stop_and_join();         // finishes what is queued: nothing writes user memory afterwards
    ctx->pipelines.clear();
    if (!ctx->devices.empty()) {
        cudaSetDevice(ctx->devices[0]);
        cudaDeviceSynchronize();
    }
    for (const sn_ctx::Pinned& r : ctx->pinned) cudaHostUnregister(r.base);
    ctx->dev_state.release();
    for (DevBuf& c : ctx->carry) c.release();
    for (int i = 0; i < kTaskRing; ++i) {
        ctx->dev_tasks[i].release();
        ctx->dev_tasks_host[i].release();
        if (ctx->dev_task_free[i]) cudaEventDestroy(ctx->dev_task_free[i]);
    }
    if (ctx->own_compute) cudaStreamDestroy(ctx->own_compute);
    cudaGetLastError();
    delete ctx;
}

int sangnom_cuda_get_stats(sn_ctx* ctx, sn_stats* out)
{
    if (!ctx || !out) return SN_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->stats_mu);
    *out = ctx->stats;
    return SN_OK;
}

void sangnom_cuda_reset_stats(sn_ctx* ctx)
{
    if (!ctx) return;
    std::lock_guard<std::mutex> lk(ctx->stats_mu);
    ctx->stats = sn_stats{};
}

int sangnom_cuda_synchronize(sn_ctx* ctx)
{
    if (!ctx) return SN_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->dev_mu);
    SN_CUDA(ctx, cudaSetDevice(ctx->devices[0]));
    SN_CUDA(ctx, cudaStreamSynchronize(ctx->own_compute));
    return SN_OK;
}

// ---------------------------------------------------------------------------------------------
static int process_planes_device_impl(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, void* cuda_stream)
{
    if (!ctx) return SN_ERR_INVALID;
    if (njobs < 0 || (njobs > 0 && !jobs)) return ctx->fail(SN_ERR_INVALID, "bad job list");
    if (njobs == 0) return SN_OK;
    std::lock_guard<std::mutex> lk(ctx->dev_mu);
    SN_CUDA(ctx, cudaSetDevice(ctx->devices[0]));
    cudaStream_t stream = cuda_stream == SN_STREAM_CONTEXT ? ctx->own_compute : static_cast<cudaStream_t>(cuda_stream);
    const int sb = ctx->sample_bytes;

    std::vector<FramePlan> frames;
    int rc = plan_frames(ctx, jobs, njobs, true, frames);
    if (rc != SN_OK) return rc;

    size_t state_total = 0, ntasks = 0;
    for (FramePlan& f : frames) { f.state_off = state_total; state_total += align_up(f.state_bytes, 256); ntasks += f.passes.size(); }
    // The scratch is shared by all device calls of this context: they must be stream-ordered.
    if (state_total > ctx->dev_state.bytes) {
        SN_CUDA(ctx, cudaStreamSynchronize(stream));
        SN_CUDA(ctx, ctx->dev_state.ensure(state_total));
    }

    // whole-plane copies of unprocessed planes; processed planes need no copy at all: the kernel reads the
    // kept rows where they are and writes them into dst itself when they are not already there
    for (FramePlan& f : frames) {
        for (const sn_plane_job* c : f.copies) {
            if (c->src == c->dst) continue;
            SN_CUDA(ctx, cudaMemcpy2DAsync(c->dst, (size_t)c->dst_pitch, c->src, (size_t)c->src_pitch, (size_t)c->width * sb,
                                           (size_t)c->dst_height, cudaMemcpyDeviceToDevice, stream));
        }
        place_state(ctx, f, static_cast<char*>(ctx->dev_state.p) + f.state_off);
    }

    std::vector<std::vector<sn::PlaneTask>> by_pass;
    for (FramePlan& f : frames)
        for (size_t q = 0; q < f.passes.size(); ++q) {
            const Pass& p = f.passes[q];
            const sn_plane_job& jb = *p.job;
            const char* kept0;
            size_t step;
            if (jb.mode == SN_MODE_INPLACE) { kept0 = static_cast<const char*>(jb.dst) + (ptrdiff_t)jb.offset * jb.dst_pitch; step = 2 * (size_t)jb.dst_pitch; }
            else if (jb.mode == SN_MODE_FIELD) { kept0 = static_cast<const char*>(jb.src) + (ptrdiff_t)jb.offset * jb.src_pitch; step = 2 * (size_t)jb.src_pitch; }
            else { kept0 = static_cast<const char*>(jb.src); step = (size_t)jb.src_pitch; }
            add_task(ctx, by_pass, q, make_task(ctx, p, jb.dst, (size_t)jb.dst_pitch, kept0, step, -1));
        }

    // Task arrays travel through a small ring of pinned/device buffers; all entries are grown
    // together (one stream sync, first call only) so steady-state submission never blocks.
    const size_t task_bytes = ntasks * sizeof(sn::PlaneTask);
    if (task_bytes > ctx->dev_tasks[0].bytes) {
        SN_CUDA(ctx, cudaStreamSynchronize(stream));
        for (int i = 0; i < kTaskRing; ++i) {
            SN_CUDA(ctx, cudaEventSynchronize(ctx->dev_task_free[i]));
            SN_CUDA(ctx, ctx->dev_tasks_host[i].ensure(task_bytes));
            SN_CUDA(ctx, ctx->dev_tasks[i].ensure(task_bytes));
        }
    }
    const int slot = ctx->dev_ring_pos;
    ctx->dev_ring_pos = (ctx->dev_ring_pos + 1) % kTaskRing;
    SN_CUDA(ctx, cudaEventSynchronize(ctx->dev_task_free[slot]));     // ring entry no longer read by an earlier upload
    SN_CUDA(ctx, upload_tasks(by_pass, static_cast<sn::PlaneTask*>(ctx->dev_tasks_host[slot].p), static_cast<sn::PlaneTask*>(ctx->dev_tasks[slot].p), stream));
    SN_CUDA(ctx, launch_passes(ctx, by_pass, static_cast<sn::PlaneTask*>(ctx->dev_tasks[slot].p), stream));
    SN_CUDA(ctx, cudaEventRecord(ctx->dev_task_free[slot], stream));
    {
        std::lock_guard<std::mutex> sl(ctx->stats_mu);
        ctx->stats.frames += frames.size();
    }
    return SN_OK;
}

// ---------------------------------------------------------------------------------------------
// Host entry, front end: plan, classify the buffers, cut into chunks, deal them to the pipelines.
static int submit_impl(sn_ctx* ctx, const sn_plane_job* user_jobs, int njobs, sn_ticket* ticket)
{
    if (!ctx) return SN_ERR_INVALID;
    if (!ticket) return ctx->fail(SN_ERR_INVALID, "null ticket pointer");
    if (njobs < 0 || (njobs > 0 && !user_jobs)) return ctx->fail(SN_ERR_INVALID, "bad job list");
    std::unique_ptr<Batch> batch(new Batch());
    batch->jobs.assign(user_jobs, user_jobs + njobs);
    int rc = plan_frames(ctx, batch->jobs.data(), njobs, false, batch->frames);
    if (rc != SN_OK) return rc;
    // pinned or pageable, and which pinned allocation: decides how every plane travels (sangnom_ctx.h, Pass::Xfer)
    if (njobs > 0) {
        cudaSetDevice(ctx->devices[0]);
        PinnedLookup pinned;
        for (FramePlan& f : batch->frames)
            for (Pass& p : f.passes) {
                p.src_pinned = pinned.find(p.job->src);
                p.dst_pinned = pinned.find(p.job->dst);
            }
    }
    const size_t nframes = batch->frames.size();
    const size_t chunk_frames = std::max<size_t>(1, (size_t)ctx->frames_in_flight / kSlots);
    const size_t nchunks = (nframes + chunk_frames - 1) / chunk_frames;
    batch->chunks_left.store((int)nchunks);

    std::lock_guard<std::mutex> lk(ctx->mu);
    batch->ticket = ++ctx->last_ticket;
    *ticket = batch->ticket;
    Batch* const b = batch.get();
    ctx->batches.push_back(std::move(batch));
    sn::plan_chunks(nframes, chunk_frames, (int)ctx->pipelines.size(), ctx->next_pipeline,
                    [&](const sn::ChunkDeal& d) { ctx->pipelines[(size_t)d.pipeline]->push(Chunk{ b, d.first, d.last }); });
    return SN_OK;
}

// Sleep until every batch up to `ticket` has finished; report the first error among them and forget them.
static int wait_impl(sn_ctx* ctx, sn_ticket ticket)
{
    if (!ctx) return SN_ERR_INVALID;
    std::unique_lock<std::mutex> lk(ctx->mu);
    if (ticket > ctx->last_ticket) { lk.unlock(); return ctx->fail(SN_ERR_INVALID, "unknown ticket"); }
    ctx->done_cv.wait(lk, [&] {
        for (const auto& b : ctx->batches) {
            if (b->ticket > ticket) break;
            if (b->chunks_left.load(std::memory_order_acquire) != 0) return false;
        }
        return true;
    });
    int status = SN_OK;
    std::string msg;
    while (!ctx->batches.empty() && ctx->batches.front()->ticket <= ticket) {
        Batch& b = *ctx->batches.front();
        if (status == SN_OK && b.status != SN_OK) { status = b.status; msg = b.error; }
        ctx->batches.pop_front();
    }
    lk.unlock();
    if (status != SN_OK) return ctx->fail(status, "%s", msg.c_str());
    return SN_OK;
}

static int process_planes_impl(sn_ctx* ctx, const sn_plane_job* jobs, int njobs)
{
    if (!ctx) return SN_ERR_INVALID;
    if (njobs == 0) return SN_OK;
    sn_ticket t = 0;
    const int rc = submit_impl(ctx, jobs, njobs, &t);
    return rc == SN_OK ? wait_impl(ctx, t) : rc;
}

// The C boundary does not let C++ exceptions through: allocation failures inside the library become SN_ERR_NOMEM.
#define SN_NOTHROW(ctx, call)                                                                   \
    try { return (call); }                                                                      \
    catch (const std::bad_alloc&) { if (ctx) (ctx)->fail(SN_ERR_NOMEM, "out of host memory"); return SN_ERR_NOMEM; } \
    catch (const std::exception& ex) { if (ctx) (ctx)->fail(SN_ERR_INVALID, "%s", ex.what()); return SN_ERR_INVALID; } \
    catch (...) { if (ctx) (ctx)->fail(SN_ERR_INVALID, "unknown C++ exception"); return SN_ERR_INVALID; }

int sangnom_cuda_process_planes_device(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, void* cuda_stream)
{
    SN_NOTHROW(ctx, process_planes_device_impl(ctx, jobs, njobs, cuda_stream))
}
int sangnom_cuda_submit(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, sn_ticket* ticket) { SN_NOTHROW(ctx, submit_impl(ctx, jobs, njobs, ticket)) }
int sangnom_cuda_wait(sn_ctx* ctx, sn_ticket ticket) { SN_NOTHROW(ctx, wait_impl(ctx, ticket)) }
int sangnom_cuda_process_planes(sn_ctx* ctx, const sn_plane_job* jobs, int njobs) { SN_NOTHROW(ctx, process_planes_impl(ctx, jobs, njobs)) }

}  // extern "C"
