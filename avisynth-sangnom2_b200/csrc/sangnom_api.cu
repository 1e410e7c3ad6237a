// libsangnom_cuda C ABI (include/sangnom_cuda.h): context, streams, staging, frame planning.
//
// Host-side counterpart of the reference's GetFrame plane loop (/root/reference/src/
// SangNom2.cpp:346-394) for a BATCH of frames: kept-field upload (the BitBlt at :361-377 becomes
// a strided H2D copy straight into the device plane), the per-plane `process` call (:393) becomes
// one thread block of the row-sweep kernel, and the finished planes are copied back.
// There is no CPU compute path in this file: without a CUDA device every call fails.
#include "sangnom_cuda.h"
#include "sangnom_kernels.h"
#include "sangnom_plan.h"
#include "host_copy_pool.h"

using sn_host::CopyPool;
using sn_host::RowCopy;

#include <atomic>
#include <condition_variable>
#include <functional>
#include <thread>

#include <cuda.h>      // driver API types only; the entry point is looked up at run time (no libcuda link dependency)

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include <chrono>

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        need = (need + 0xFFFFF) & ~(size_t)0xFFFFF;   // 1 MiB granules
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        need = (need + 0xFFFFF) & ~(size_t)0xFFFFF;
        cudaError_t e = cudaHostAlloc(&p, need, cudaHostAllocDefault);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
};

// A processed plane of one frame, after validation.
struct Pass {
    const sn_plane_job* job;
    int W, H, n;             // samples, rows, kept rows
    int R;                   // pool rows to sweep
    int cone;                // dependency-cone bound (sangnom_plan.h)
    sn::CostState in{}, out{};
    // host path: how the kept field gets up and the finished plane gets down.
    //   STAGED  pageable host memory: CPU packs rows into the slot's pinned staging, one contiguous DMA
    //   LINEAR  pinned, rows contiguous and 16-byte multiples: one contiguous DMA straight from/to user memory
    //           (upload: the whole source plane - a strided kept-field copy runs ~8x slower on this DMA engine)
    //   PITCHED pinned, anything else: 2-D DMA
    enum Xfer { STAGED, LINEAR, PITCHED };
    Xfer up = STAGED, down = STAGED;
    size_t src_off = 0, src_bytes = 0, src_pitch = 0;   // device copy of the source rows (bytes)
    size_t src_first = 0, src_step = 0;                 // kept row 0 and kept-row step inside it (bytes)
    size_t dst_off = 0, dst_pitch = 0;                  // device dst plane
    size_t stage_in_off = 0, stage_out_off = 0;
    int src_pinned = 0, dst_pinned = 0;                // 0 = pageable, else 1 + pinned allocation index (PinnedLookup)
};

struct FramePlan {
    int key;
    std::vector<Pass> passes;            // processed planes in plane order (at most 3)
    std::vector<const sn_plane_job*> copies;
    size_t state_bytes = 0;              // cost-state scratch of this frame
    size_t state_off = 0;
};

// One pipeline slot: a chunk of frames resident on the device.
struct Slot {
    DevBuf planes, state, tasks;
    PinnedBuf tasks_host, stage_in, stage_out;
    cudaStream_t compute = nullptr;
    cudaEvent_t h2d_start = nullptr, h2d_done = nullptr, k_start = nullptr, kernels_done = nullptr, d2h_start = nullptr, d2h_done = nullptr;
    bool busy = false;
    std::vector<FramePlan*> frames;      // frames of the chunk in flight (for the pageable copy-out)
    uint64_t ticket = 0;                 // batch the chunk belongs to
};

// One submitted job list: the jobs are copied so that the plans may outlive the caller's array.
struct Batch {
    uint64_t ticket = 0;
    std::vector<sn_plane_job> jobs;
    std::vector<FramePlan> frames;
};

constexpr int kSlots = 4;
constexpr int kTaskRing = 8;

}  // namespace

struct sn_ctx {
    sn_config cfg{};
    int sample_bytes = 1;
    int S = 0, Hb = 0;
    int frames_in_flight = 0;
    cudaStream_t h2d = nullptr, d2h = nullptr, own_compute = nullptr;
    cudaEvent_t trace_base = nullptr;
    bool trace_base_set = false;
    Slot slots[kSlots];
    int next_slot = 0;                   // slots are used round-robin, so this is also the oldest one in flight
    uint64_t last_ticket = 0;
    std::deque<std::unique_ptr<Batch>> batches;
    // device-entry resources
    DevBuf dev_state;
    // persistent-pool mode (SN_FLAG_PERSISTENT_POOL): the pool state between frames, ping-pong
    bool persistent = false;
    bool saturate = false;               // SN_FLAG_SATURATE: the reference's SSE2 (opt=1) arithmetic flavour
    DevBuf carry[2];
    int carry_pos = 0;                   // carry[carry_pos] holds the state the next frame starts from
    DevBuf dev_tasks[kTaskRing];
    PinnedBuf dev_tasks_host[kTaskRing];
    cudaEvent_t dev_task_free[kTaskRing] = {};
    int dev_ring_pos = 0;
    sn_stats stats{};
    std::string error;
    std::mutex mu;
    std::unique_ptr<CopyPool> copy_pool;   // created on the first staged (pageable) transfer
    CopyPool& pool()
    {
        if (!copy_pool) {
            const char* v = getenv("SANGNOM_B200_COPY_THREADS");
            int n = v && *v ? atoi(v) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
            copy_pool.reset(new CopyPool(std::max(0, std::min(n, 64) - 1)));      // the caller is one of the n
        }
        return *copy_pool;
    }

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
    int cuda_fail(cudaError_t e, const char* what)
    {
        return fail(SN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
};

#define SN_CUDA(ctx, call)                                          \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, #call); \
    } while (0)

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Which pinned allocation a host pointer lies in. DMA transfers may only be merged inside one allocation (a copy that
// spans two cudaHostAlloc blocks fails even when they are neighbours in the address space), so pinned-ness and the
// allocation's extent are looked up together; lookups are cached per submitted batch, one driver query per allocation.
struct HostRange { uintptr_t lo = 0, hi = 0; };

class PinnedLookup {
    using GetAttr = CUresult (*)(void*, CUpointer_attribute, CUdeviceptr);
    GetAttr get_ = nullptr;
    std::vector<HostRange> known_;
public:
    static GetAttr driver_entry()          // looked up once per process
    {
        static const GetAttr fn = [] {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
                return reinterpret_cast<GetAttr>(p);
            cudaGetLastError();
            return static_cast<GetAttr>(nullptr);
        }();
        return fn;
    }
    PinnedLookup() : get_(driver_entry()) {}
    // 0 = pageable; otherwise 1 + index of the allocation (stable for this object's lifetime)
    int find(const void* p)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        for (size_t i = 0; i < known_.size(); ++i)
            if (a >= known_[i].lo && a < known_[i].hi) return (int)i + 1;
        HostRange r;
        if (get_) {
            CUmemorytype type{};
            CUdeviceptr base = 0;
            size_t size = 0;
            if (get_(&type, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)a) != CUDA_SUCCESS || type != CU_MEMORYTYPE_HOST) return 0;
            if (get_(&base, CU_POINTER_ATTRIBUTE_RANGE_START_ADDR, (CUdeviceptr)a) != CUDA_SUCCESS ||
                get_(&size, CU_POINTER_ATTRIBUTE_RANGE_SIZE, (CUdeviceptr)a) != CUDA_SUCCESS || size == 0) {
                r.lo = a; r.hi = a + 1;            // pinned, extent unknown: never merged with anything
            } else {
                r.lo = (uintptr_t)base; r.hi = r.lo + size;
            }
        } else {
            cudaPointerAttributes at{};
            if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 0; }
            if (at.type != cudaMemoryTypeHost) return 0;
            r.lo = a; r.hi = a + 1;
        }
        known_.push_back(r);
        return (int)known_.size();
    }
};

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
        if (jb.dst_pitch % sb != 0) return ctx->fail(SN_ERR_INVALID, "job %d: pitch not a multiple of the sample size", j);
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
        for (int q = 0; q < m; ++q) { f.passes[q].R = geo[q].sweep_rows; f.passes[q].cone = geo[q].cone; f.passes[q].in = geo[q].in; f.passes[q].out = geo[q].out; }
    }
    return SN_OK;
}

// Resolve the frame's cost-state regions inside its scratch; in persistent-pool mode also hook the frame into the
// chain of pool states (frames are placed in submission order, which is the order they run in).
void place_state(sn_ctx* ctx, FramePlan& f, char* base)
{
    for (Pass& p : f.passes) { sn::plan_place_state(p.in, base); sn::plan_place_state(p.out, base); }
    if (ctx->persistent && !f.passes.empty() && ctx->carry[0].p) {
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

sn::PlaneTask make_task(const sn_ctx* ctx, const Pass& p, void* plane, size_t pitch_bytes, const void* kept0, size_t kept_step_bytes)
{
    sn::PlaneTask t{};
    t.plane = plane;
    t.pitch = (long long)(pitch_bytes / ctx->sample_bytes);
    t.src = kept0;
    t.src_pitch = (long long)(kept_step_bytes / ctx->sample_bytes);
    // in place when the kept rows already sit at rows offset, offset+2, .. of the dst plane
    t.copy_kept = !(kept0 == static_cast<char*>(plane) + (size_t)p.job->offset * pitch_bytes && kept_step_bytes == 2 * pitch_bytes);
    t.width = p.W; t.height = p.H; t.offset = p.job->offset;
    t.kept_rows = p.n; t.sweep_rows = p.R; t.cone = p.cone;
    t.thr_f = p.job->threshold;
    // `const T aaf` (reference SangNom2.cpp:162,272): float -> T for integer samples. Truncate
    // toward zero, then wrap to the container width - what x86-64 does for the (undefined in C++)
    // negative case the legacy SangNom() entry can produce.
    if (ctx->sample_bytes == 1) t.thr_i = (int)p.job->threshold & 0xFF;
    else if (ctx->sample_bytes == 2) t.thr_i = (int)p.job->threshold & 0xFFFF;
    t.in = p.in; t.out = p.out;
    return t;
}

void host_copy_plane(const sn_plane_job& jb, int sb, std::vector<RowCopy>& copies)
{
    if (jb.src == jb.dst) return;
    copies.push_back(RowCopy{ static_cast<char*>(jb.dst), static_cast<const char*>(jb.src), jb.dst_pitch, jb.src_pitch, (size_t)jb.width * sb, jb.dst_height });
}

// Launch the passes of a set of frames: one kernel per pass index (all first planes, then all
// second planes, ...), stream-ordered so pass q+1 of a frame sees pass q's cost state.
int upload_tasks(sn_ctx* ctx, const std::vector<std::vector<sn::PlaneTask>>& by_pass, sn::PlaneTask* host_tasks,
                 sn::PlaneTask* dev_tasks, cudaStream_t stream)
{
    size_t total = 0;
    for (auto& v : by_pass) { std::memcpy(host_tasks + total, v.data(), v.size() * sizeof(sn::PlaneTask)); total += v.size(); }
    if (total == 0) return SN_OK;
    SN_CUDA(ctx, cudaMemcpyAsync(dev_tasks, host_tasks, total * sizeof(sn::PlaneTask), cudaMemcpyHostToDevice, stream));
    return SN_OK;
}

int launch_passes(sn_ctx* ctx, const std::vector<std::vector<sn::PlaneTask>>& by_pass, sn::PlaneTask* dev_tasks, cudaStream_t stream)
{
    size_t pos = 0;
    for (auto& v : by_pass) {
        if (v.empty()) continue;
        bool narrow = false;                 // any plane of this launch narrower than the pool by a thread's 8 columns or more
        for (const sn::PlaneTask& t : v) narrow = narrow || t.width + 8 <= ctx->S;
        SN_CUDA(ctx, sn::launch_plane_tasks(ctx->sample_bytes, dev_tasks + pos, (int)v.size(), sn::make_geometry(ctx->S, ctx->Hb, ctx->saturate, narrow), stream));
        ctx->stats.kernel_launches += 1;
        ctx->stats.planes_processed += v.size();
        pos += v.size();
    }
    return SN_OK;
}

// A run of bytes that is contiguous both in (pinned) host memory and in the device slot: one DMA transfer.
struct Segment { char* host; char* dev; size_t bytes; int alloc; };

// alloc: which pinned allocation `host` lies in (runs are merged only inside one allocation)
void add_segment(std::vector<Segment>& v, void* host, void* dev, size_t bytes, int alloc)
{
    char* h = static_cast<char*>(host);
    char* d = static_cast<char*>(dev);
    if (!v.empty() && v.back().alloc == alloc && v.back().host + v.back().bytes == h && v.back().dev + v.back().bytes == d) v.back().bytes += bytes;
    else v.push_back(Segment{ h, d, bytes, alloc });
}

cudaError_t flush_segments(std::vector<Segment>& v, cudaMemcpyKind kind, cudaStream_t stream)
{
    cudaError_t e = cudaSuccess;
    for (const Segment& g : v) {
        e = kind == cudaMemcpyHostToDevice ? cudaMemcpyAsync(g.dev, g.host, g.bytes, kind, stream) : cudaMemcpyAsync(g.host, g.dev, g.bytes, kind, stream);
        if (e != cudaSuccess) break;
    }
    v.clear();
    return e;
}

// SANGNOM_UPLOAD=field: upload only the kept rows of pinned contiguous planes (2-D DMA) instead of the whole plane.
const bool g_upload_kept_only = [] { const char* v = getenv("SANGNOM_UPLOAD"); return v && std::strcmp(v, "field") == 0; }();

// Copy the finished planes of a chunk from pinned staging to pageable destinations.
int drain_slot(sn_ctx* ctx, Slot& s)
{
    if (!s.busy) return SN_OK;
    SN_CUDA(ctx, cudaEventSynchronize(s.d2h_done));
    static const bool trace = getenv("SANGNOM_TRACE") != nullptr;
    if (trace) {
        float t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
        cudaEventElapsedTime(&t0, ctx->trace_base, s.h2d_start);
        cudaEventElapsedTime(&t1, ctx->trace_base, s.h2d_done);
        cudaEventElapsedTime(&t2, ctx->trace_base, s.k_start);
        cudaEventElapsedTime(&t3, ctx->trace_base, s.kernels_done);
        cudaEventElapsedTime(&t4, ctx->trace_base, s.d2h_start);
        cudaEventElapsedTime(&t5, ctx->trace_base, s.d2h_done);
        fprintf(stderr, "[sangnom] chunk of %3zu frames: h2d %6.2f..%6.2f  kernels %6.2f..%6.2f  d2h %6.2f..%6.2f ms\n",
                s.frames.size(), t0, t1, t2, t3, t4, t5);
    }
    const int sb = ctx->sample_bytes;
    std::vector<RowCopy> copies;
    for (FramePlan* f : s.frames)
        for (Pass& p : f->passes) {
            if (p.down != Pass::STAGED) continue;
            const sn_plane_job& jb = *p.job;
            const char* st = static_cast<const char*>(s.stage_out.p) + p.stage_out_off;
            copies.push_back(RowCopy{ static_cast<char*>(jb.dst), st, jb.dst_pitch, (ptrdiff_t)p.dst_pitch, (size_t)p.W * sb, p.H });
        }
    if (!copies.empty()) ctx->pool().run(copies);
    s.frames.clear();
    s.busy = false;
    return SN_OK;
}

}  // namespace

extern "C" {

int sangnom_cuda_abi_version(void) { return SANGNOM_CUDA_ABI_VERSION; }

const char* sangnom_cuda_last_error(sn_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

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
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void sangnom_cuda_host_free(void* p) { if (p) cudaFreeHost(p); }

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
    if (cfg->device < 0 || cfg->device >= count) { g_create_error = "device ordinal out of range"; return SN_ERR_INVALID; }
    cudaDeviceProp prop{};
    e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return SN_ERR_CUDA; }
    if (prop.major != 10) {
        char b[160];
        snprintf(b, sizeof b, "device %d is sm_%d%d; this library carries sm_100a code only", cfg->device, prop.major, prop.minor);
        g_create_error = b;
        return SN_ERR_CUDA;
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
    if (cfg->max_frames_in_flight > 0) {
        ctx->frames_in_flight = cfg->max_frames_in_flight;
    } else {
        // default: one frame per SM over the kSlots chunks in flight - chunks small enough that the pipeline's
        // ramp (first upload, last download, one kernel latency) stays short, large enough that the chunks whose
        // kernels overlap fill the GPU - capped at ~24 GB of device memory for planes + cost state (about 2.5 x
        // the three planes of a frame)
        const double per_frame = 2.5 * 3.0 * (double)S * (double)cfg->pool_height * (double)cfg->sample_type;
        const long long fit = (long long)(24.0e9 / per_frame);
        ctx->frames_in_flight = (int)std::min<long long>(prop.multiProcessorCount, std::max<long long>(fit, 12));
    }
    auto bail = [&](cudaError_t err, const char* what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        sangnom_cuda_destroy(ctx);
        return SN_ERR_CUDA;
    };
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->own_compute, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (Slot& s : ctx->slots) {
        if ((e = cudaStreamCreateWithFlags(&s.compute, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
        if ((e = cudaEventCreate(&s.h2d_start)) != cudaSuccess) return bail(e, "cudaEventCreate");
        if ((e = cudaEventCreate(&s.k_start)) != cudaSuccess) return bail(e, "cudaEventCreate");
        if ((e = cudaEventCreate(&s.d2h_start)) != cudaSuccess) return bail(e, "cudaEventCreate");
        if ((e = cudaEventCreate(&s.h2d_done)) != cudaSuccess) return bail(e, "cudaEventCreate");
        if ((e = cudaEventCreate(&s.kernels_done)) != cudaSuccess) return bail(e, "cudaEventCreate");
        if ((e = cudaEventCreate(&s.d2h_done)) != cudaSuccess) return bail(e, "cudaEventCreate");
    }
    for (int i = 0; i < kTaskRing; ++i)
        if ((e = cudaEventCreateWithFlags(&ctx->dev_task_free[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->trace_base)) != cudaSuccess) return bail(e, "cudaEventCreate");
    ctx->saturate = (cfg->flags & SN_FLAG_SATURATE) != 0;
    if (cfg->flags & SN_FLAG_PERSISTENT_POOL) {
        ctx->persistent = true;
        const size_t bytes = sn::plan_carry_bytes(S, Hb, cfg->sample_type);
        for (DevBuf& c : ctx->carry)
            if (bytes) {
                if ((e = c.ensure(bytes)) != cudaSuccess) return bail(e, "pool state allocation");
                if ((e = cudaMemset(c.p, 0, bytes)) != cudaSuccess) return bail(e, "cudaMemset");      // the zero-filled pool of a new instance
            }
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bail(e, "cudaDeviceSynchronize");
    }
    *out = ctx;
    return SN_OK;
}

void sangnom_cuda_destroy(sn_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    cudaDeviceSynchronize();
    for (Slot& s : ctx->slots) {
        s.planes.release(); s.state.release(); s.tasks.release();
        s.tasks_host.release(); s.stage_in.release(); s.stage_out.release();
        if (s.compute) cudaStreamDestroy(s.compute);
        if (s.h2d_start) cudaEventDestroy(s.h2d_start);
        if (s.k_start) cudaEventDestroy(s.k_start);
        if (s.d2h_start) cudaEventDestroy(s.d2h_start);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.kernels_done) cudaEventDestroy(s.kernels_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
    }
    ctx->dev_state.release();
    for (DevBuf& c : ctx->carry) c.release();
    for (int i = 0; i < kTaskRing; ++i) {
        ctx->dev_tasks[i].release();
        ctx->dev_tasks_host[i].release();
        if (ctx->dev_task_free[i]) cudaEventDestroy(ctx->dev_task_free[i]);
    }
    if (ctx->trace_base) cudaEventDestroy(ctx->trace_base);
    if (ctx->h2d) cudaStreamDestroy(ctx->h2d);
    if (ctx->d2h) cudaStreamDestroy(ctx->d2h);
    if (ctx->own_compute) cudaStreamDestroy(ctx->own_compute);
    cudaGetLastError();
    delete ctx;
}

int sangnom_cuda_get_stats(sn_ctx* ctx, sn_stats* out)
{
    if (!ctx || !out) return SN_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    *out = ctx->stats;
    return SN_OK;
}

void sangnom_cuda_reset_stats(sn_ctx* ctx)
{
    if (!ctx) return;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->stats = sn_stats{};
}

int sangnom_cuda_synchronize(sn_ctx* ctx)
{
    if (!ctx) return SN_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    SN_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    SN_CUDA(ctx, cudaStreamSynchronize(ctx->own_compute));
    return SN_OK;
}

// ---------------------------------------------------------------------------------------------
static int process_planes_device_impl(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, void* cuda_stream)
{
    if (!ctx) return SN_ERR_INVALID;
    if (njobs < 0 || (njobs > 0 && !jobs)) return ctx->fail(SN_ERR_INVALID, "bad job list");
    if (njobs == 0) return SN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    SN_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t stream = cuda_stream == SN_STREAM_CONTEXT ? ctx->own_compute : static_cast<cudaStream_t>(cuda_stream);
    const int sb = ctx->sample_bytes;

    static const bool trace = getenv("SANGNOM_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
    const auto t0 = now();
    std::vector<FramePlan> frames;
    int rc = plan_frames(ctx, jobs, njobs, true, frames);
    if (rc != SN_OK) return rc;
    const auto t1 = now();

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

    const auto t2 = now();
    std::vector<std::vector<sn::PlaneTask>> by_pass;
    for (FramePlan& f : frames)
        for (size_t q = 0; q < f.passes.size(); ++q)
        {
            const Pass& p = f.passes[q];
            const sn_plane_job& jb = *p.job;
            const char* kept0;
            size_t step;
            if (jb.mode == SN_MODE_INPLACE) { kept0 = static_cast<const char*>(jb.dst) + (ptrdiff_t)jb.offset * jb.dst_pitch; step = 2 * (size_t)jb.dst_pitch; }
            else if (jb.mode == SN_MODE_FIELD) { kept0 = static_cast<const char*>(jb.src) + (ptrdiff_t)jb.offset * jb.src_pitch; step = 2 * (size_t)jb.src_pitch; }
            else { kept0 = static_cast<const char*>(jb.src); step = (size_t)jb.src_pitch; }
            add_task(ctx, by_pass, q, make_task(ctx, p, jb.dst, (size_t)jb.dst_pitch, kept0, step));
        }
    const auto t3 = now();

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
    rc = upload_tasks(ctx, by_pass, static_cast<sn::PlaneTask*>(ctx->dev_tasks_host[slot].p),
                      static_cast<sn::PlaneTask*>(ctx->dev_tasks[slot].p), stream);
    if (rc != SN_OK) return rc;
    rc = launch_passes(ctx, by_pass, static_cast<sn::PlaneTask*>(ctx->dev_tasks[slot].p), stream);
    if (rc != SN_OK) return rc;
    SN_CUDA(ctx, cudaEventRecord(ctx->dev_task_free[slot], stream));
    ctx->stats.frames += frames.size();
    if (trace) {
        const auto t4 = now();
        fprintf(stderr, "[sangnom] device submit: plan %ld us, place %ld us, tasks %ld us, ring+launch %ld us (%d jobs)\n",
                us(t0, t1), us(t1, t2), us(t2, t3), us(t3, t4), njobs);
    }
    return SN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
namespace {

// Wait for every chunk of batches <= ticket (slots are drained oldest first) and forget finished batches.
int drain_through(sn_ctx* ctx, uint64_t ticket)
{
    int status = SN_OK;
    for (int k = 0; k < kSlots; ++k) {
        Slot& s = ctx->slots[(ctx->next_slot + k) % kSlots];
        if (!s.busy || s.ticket > ticket) continue;
        if (status == SN_OK) status = drain_slot(ctx, s);
        else { cudaEventSynchronize(s.d2h_done); s.frames.clear(); s.busy = false; }
    }
    while (!ctx->batches.empty() && ctx->batches.front()->ticket <= ticket) ctx->batches.pop_front();
    return status;
}

int submit_locked(sn_ctx* ctx, const sn_plane_job* user_jobs, int njobs, uint64_t* ticket_out)
{
    SN_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const int sb = ctx->sample_bytes;

    static const bool trace = getenv("SANGNOM_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    std::unique_ptr<Batch> batch(new Batch());
    batch->ticket = ++ctx->last_ticket;
    batch->jobs.assign(user_jobs, user_jobs + njobs);
    const sn_plane_job* jobs = batch->jobs.data();
    std::vector<FramePlan>& frames = batch->frames;
    const uint64_t ticket = batch->ticket;
    *ticket_out = ticket;
    int rc = plan_frames(ctx, jobs, njobs, false, frames);
    if (rc != SN_OK) return rc;
    ctx->batches.push_back(std::move(batch));
    PinnedLookup pinned;
    for (FramePlan& f : frames)
        for (Pass& p : f.passes) {
            p.src_pinned = pinned.find(p.job->src);
            p.dst_pinned = pinned.find(p.job->dst);
        }

    const auto t_planned = std::chrono::steady_clock::now();
    if (!ctx->trace_base_set) { cudaEventRecord(ctx->trace_base, ctx->h2d); ctx->trace_base_set = true; }     // one time origin per context
    const size_t chunk_frames = std::max<size_t>(1, (size_t)ctx->frames_in_flight / kSlots);
    size_t next = 0;
    int status = SN_OK;
    while (next < frames.size() && status == SN_OK) {
        Slot& s = ctx->slots[ctx->next_slot];
        ctx->next_slot = (ctx->next_slot + 1) % kSlots;
        if ((status = drain_slot(ctx, s)) != SN_OK) break;     // also guarantees the slot's device buffers are free

        const size_t first = next, last = std::min(frames.size(), next + chunk_frames);
        next = last;

        // placement inside the slot: one region for the uploaded source rows and one for the finished planes,
        // each filled in job order. LINEAR planes are packed at 16-byte granules (bulk copies and vector accesses
        // need no more), so planes that are contiguous in host memory are contiguous on the device as well and their
        // DMA transfers merge into one (a 518 KB chroma plane copied alone runs at ~39 GB/s, a merged run at ~53).
        size_t src_bytes_total = 0, dst_bytes_total = 0, state_bytes = 0, in_bytes = 0, out_bytes = 0, ntasks = 0;
        for (size_t k = first; k < last; ++k) {
            FramePlan& f = frames[k];
            f.state_off = state_bytes;
            state_bytes += align_up(f.state_bytes, 256);
            for (Pass& p : f.passes) {
                const sn_plane_job& jb = *p.job;
                const size_t row = (size_t)p.W * sb, rowpad = align_up(row, 16);
                const int src_rows = jb.mode == SN_MODE_FIELD ? p.H : p.n;
                const bool field = jb.mode == SN_MODE_FIELD;
                if (!p.src_pinned) {                                   // kept rows packed by the CPU
                    p.up = Pass::STAGED; p.src_pitch = rowpad; p.src_bytes = rowpad * p.n; p.src_first = 0; p.src_step = rowpad;
                    p.stage_in_off = in_bytes; in_bytes += p.src_bytes;
                } else if ((size_t)jb.src_pitch == row && row % 16 == 0 && !g_upload_kept_only) {   // whole source plane, contiguous DMA
                    p.up = Pass::LINEAR; p.src_pitch = row; p.src_bytes = row * src_rows;
                    p.src_first = field ? (size_t)jb.offset * row : 0; p.src_step = field ? 2 * row : row;
                } else {                                               // kept rows by 2-D DMA
                    p.up = Pass::PITCHED; p.src_pitch = rowpad; p.src_bytes = rowpad * p.n; p.src_first = 0; p.src_step = rowpad;
                }
                p.src_off = src_bytes_total; src_bytes_total += p.src_bytes;          // multiples of 16 in every class
                if (!p.dst_pinned) { p.down = Pass::STAGED; p.dst_pitch = rowpad; p.stage_out_off = out_bytes; out_bytes += rowpad * p.H; }
                else if ((size_t)jb.dst_pitch == row && row % 16 == 0) { p.down = Pass::LINEAR; p.dst_pitch = row; }
                else { p.down = Pass::PITCHED; p.dst_pitch = rowpad; }
                p.dst_off = dst_bytes_total; dst_bytes_total += p.dst_pitch * p.H;
                ++ntasks;
            }
        }
        const size_t dst_base = align_up(src_bytes_total, 256);
        const size_t plane_bytes = dst_base + dst_bytes_total;
        for (size_t k = first; k < last; ++k)
            for (Pass& p : frames[k].passes) p.dst_off += dst_base;
        cudaError_t e;
        if ((e = s.planes.ensure(plane_bytes)) != cudaSuccess || (e = s.state.ensure(state_bytes)) != cudaSuccess ||
            (e = s.tasks.ensure(ntasks * sizeof(sn::PlaneTask))) != cudaSuccess ||
            (e = s.tasks_host.ensure(ntasks * sizeof(sn::PlaneTask))) != cudaSuccess ||
            (e = s.stage_in.ensure(in_bytes)) != cudaSuccess || (e = s.stage_out.ensure(out_bytes)) != cudaSuccess) {
            status = ctx->cuda_fail(e, "slot allocation");
            break;
        }

        // ---- host-side copies of the chunk, all at once on the copy pool: kept rows of pageable sources into the
        // pinned staging buffer, and the planes that are only copied (disabled planes, alpha) ----
        {
            std::vector<RowCopy> host_copies;
            for (size_t k = first; k < last; ++k) {
                FramePlan& f = frames[k];
                for (const sn_plane_job* c : f.copies) host_copy_plane(*c, sb, host_copies);
                for (Pass& p : f.passes) {
                    if (p.up != Pass::STAGED) continue;
                    const sn_plane_job& jb = *p.job;
                    const char* kept = static_cast<const char*>(jb.src) + (jb.mode == SN_MODE_FIELD ? (ptrdiff_t)jb.offset * jb.src_pitch : 0);
                    const ptrdiff_t kept_step = jb.src_pitch * (jb.mode == SN_MODE_FIELD ? 2 : 1);
                    host_copies.push_back(RowCopy{ static_cast<char*>(s.stage_in.p) + p.stage_in_off, kept, (ptrdiff_t)p.src_pitch, kept_step, (size_t)p.W * sb, p.n });
                }
            }
            if (!host_copies.empty()) ctx->pool().run(host_copies);
        }

        cudaEventRecord(s.h2d_start, ctx->h2d);
        // ---- upload (the reference's kept-field BitBlt, SangNom2.cpp:361-377, becomes DMA + the kernel's own reads) ----
        std::vector<std::vector<sn::PlaneTask>> by_pass;
        std::vector<Segment> up_segs, down_segs;
        for (size_t k = first; k < last && status == SN_OK; ++k) {
            FramePlan& f = frames[k];
            place_state(ctx, f, static_cast<char*>(s.state.p) + f.state_off);
            for (size_t q = 0; q < f.passes.size(); ++q) {
                Pass& p = f.passes[q];
                const sn_plane_job& jb = *p.job;
                const size_t row = (size_t)p.W * sb;
                const char* kept = static_cast<const char*>(jb.src) + (jb.mode == SN_MODE_FIELD ? (ptrdiff_t)jb.offset * jb.src_pitch : 0);
                const size_t kept_step = (size_t)jb.src_pitch * (jb.mode == SN_MODE_FIELD ? 2 : 1);
                char* dsrc = static_cast<char*>(s.planes.p) + p.src_off;
                e = cudaSuccess;
                if (p.up == Pass::STAGED) {
                    add_segment(up_segs, static_cast<char*>(s.stage_in.p) + p.stage_in_off, dsrc, p.src_bytes, -1);                         // the slot's own staging buffer
                } else if (p.up == Pass::LINEAR) {
                    add_segment(up_segs, const_cast<void*>(jb.src), dsrc, p.src_bytes, p.src_pinned);
                } else {
                    e = flush_segments(up_segs, cudaMemcpyHostToDevice, ctx->h2d);          // keep submission order
                    if (e == cudaSuccess)
                        e = cudaMemcpy2DAsync(dsrc, p.src_pitch, kept, kept_step, row, (size_t)p.n, cudaMemcpyHostToDevice, ctx->h2d);
                }
                if (e != cudaSuccess) { status = ctx->cuda_fail(e, "H2D copy"); break; }
                ctx->stats.h2d_bytes += p.up == Pass::PITCHED ? row * p.n : p.src_bytes;
                add_task(ctx, by_pass, q, make_task(ctx, p, static_cast<char*>(s.planes.p) + p.dst_off, p.dst_pitch, dsrc + p.src_first, p.src_step));
            }
        }
        if (status == SN_OK && (e = flush_segments(up_segs, cudaMemcpyHostToDevice, ctx->h2d)) != cudaSuccess) status = ctx->cuda_fail(e, "H2D copy");
        if (status != SN_OK) break;
        // The task array rides the upload stream too: a small copy on the compute stream would queue on the
        // same DMA engine behind the NEXT chunks' bulk uploads and hold this chunk's kernels back.
        if ((status = upload_tasks(ctx, by_pass, static_cast<sn::PlaneTask*>(s.tasks_host.p), static_cast<sn::PlaneTask*>(s.tasks.p), ctx->h2d)) != SN_OK) break;
        // persistent pool: every chunk's kernels on ONE stream, so that frames run in submission order across chunks
        const cudaStream_t compute = ctx->persistent ? ctx->own_compute : s.compute;
        if ((e = cudaEventRecord(s.h2d_done, ctx->h2d)) != cudaSuccess || (e = cudaStreamWaitEvent(compute, s.h2d_done, 0)) != cudaSuccess) {
            status = ctx->cuda_fail(e, "event"); break;
        }

        // ---- kernels ----
        cudaEventRecord(s.k_start, compute);
        status = launch_passes(ctx, by_pass, static_cast<sn::PlaneTask*>(s.tasks.p), compute);
        if (status != SN_OK) break;
        if ((e = cudaEventRecord(s.kernels_done, compute)) != cudaSuccess || (e = cudaStreamWaitEvent(ctx->d2h, s.kernels_done, 0)) != cudaSuccess) {
            status = ctx->cuda_fail(e, "event"); break;
        }

        // ---- download ----
        cudaEventRecord(s.d2h_start, ctx->d2h);
        for (size_t k = first; k < last && status == SN_OK; ++k) {
            FramePlan& f = frames[k];
            for (Pass& p : f.passes) {
                const sn_plane_job& jb = *p.job;
                const size_t row = (size_t)p.W * sb;
                const char* dplane = static_cast<char*>(s.planes.p) + p.dst_off;
                char* dplane_w = const_cast<char*>(dplane);
                e = cudaSuccess;
                if (p.down == Pass::STAGED)
                    add_segment(down_segs, static_cast<char*>(s.stage_out.p) + p.stage_out_off, dplane_w, p.dst_pitch * p.H, -1);
                else if (p.down == Pass::LINEAR)
                    add_segment(down_segs, jb.dst, dplane_w, row * p.H, p.dst_pinned);
                else {
                    e = flush_segments(down_segs, cudaMemcpyDeviceToHost, ctx->d2h);
                    if (e == cudaSuccess)
                        e = cudaMemcpy2DAsync(jb.dst, (size_t)jb.dst_pitch, dplane, p.dst_pitch, row, (size_t)p.H, cudaMemcpyDeviceToHost, ctx->d2h);
                }
                if (e != cudaSuccess) { status = ctx->cuda_fail(e, "D2H copy"); break; }
                ctx->stats.d2h_bytes += p.down == Pass::PITCHED ? row * p.H : p.dst_pitch * p.H;
            }
            s.frames.push_back(&f);
        }
        if (status == SN_OK && (e = flush_segments(down_segs, cudaMemcpyDeviceToHost, ctx->d2h)) != cudaSuccess) status = ctx->cuda_fail(e, "D2H copy");
        if (status != SN_OK) break;
        if ((e = cudaEventRecord(s.d2h_done, ctx->d2h)) != cudaSuccess) { status = ctx->cuda_fail(e, "event"); break; }
        s.busy = true;
        s.ticket = ticket;
        ctx->stats.frames += last - first;
    }
    if (trace) {
        const auto t_end = std::chrono::steady_clock::now();
        auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
        fprintf(stderr, "[sangnom] submit of %d jobs: plan+classify %ld us, chunks (incl. waiting for slots) %ld us\n", njobs, us(t_begin, t_planned), us(t_planned, t_end));
    }
    if (status != SN_OK) {
        // leave nothing in flight that writes user memory, then forget every batch
        for (int k = 0; k < kSlots; ++k) {
            Slot& s = ctx->slots[(ctx->next_slot + k) % kSlots];
            if (s.busy) { cudaEventSynchronize(s.d2h_done); s.frames.clear(); s.busy = false; }
        }
        cudaStreamSynchronize(ctx->h2d); cudaStreamSynchronize(ctx->d2h); cudaGetLastError();
        ctx->batches.clear();
    }
    return status;
}

}  // namespace

extern "C" {

static int submit_impl(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, sn_ticket* ticket)
{
    if (!ctx) return SN_ERR_INVALID;
    if (!ticket) return ctx->fail(SN_ERR_INVALID, "null ticket pointer");
    if (njobs < 0 || (njobs > 0 && !jobs)) return ctx->fail(SN_ERR_INVALID, "bad job list");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (njobs == 0) { *ticket = ctx->last_ticket; return SN_OK; }
    uint64_t t = 0;
    const int rc = submit_locked(ctx, jobs, njobs, &t);
    *ticket = t;
    return rc;
}

static int wait_impl(sn_ctx* ctx, sn_ticket ticket)
{
    if (!ctx) return SN_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ticket > ctx->last_ticket) return ctx->fail(SN_ERR_INVALID, "unknown ticket");
    SN_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    return drain_through(ctx, ticket);
}

static int process_planes_impl(sn_ctx* ctx, const sn_plane_job* jobs, int njobs)
{
    if (!ctx) return SN_ERR_INVALID;
    if (njobs < 0 || (njobs > 0 && !jobs)) return ctx->fail(SN_ERR_INVALID, "bad job list");
    if (njobs == 0) return SN_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    uint64_t t = 0;
    int rc = submit_locked(ctx, jobs, njobs, &t);
    if (rc == SN_OK) rc = drain_through(ctx, t);
    return rc;
}


// The C boundary does not let C++ exceptions through: allocation failures inside the library become SN_ERR_NOMEM.
#define SN_NOTHROW(ctx, call)                                                                   \
    try { return (call); }                                                                      \
    catch (const std::bad_alloc&) { if (ctx) (ctx)->error = "out of host memory"; return SN_ERR_NOMEM; } \
    catch (const std::exception& ex) { if (ctx) (ctx)->error = ex.what(); return SN_ERR_INVALID; } \
    catch (...) { if (ctx) (ctx)->error = "unknown C++ exception"; return SN_ERR_INVALID; }

int sangnom_cuda_process_planes_device(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, void* cuda_stream)
{
    SN_NOTHROW(ctx, process_planes_device_impl(ctx, jobs, njobs, cuda_stream))
}
int sangnom_cuda_submit(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, sn_ticket* ticket) { SN_NOTHROW(ctx, submit_impl(ctx, jobs, njobs, ticket)) }
int sangnom_cuda_wait(sn_ctx* ctx, sn_ticket ticket) { SN_NOTHROW(ctx, wait_impl(ctx, ticket)) }
int sangnom_cuda_process_planes(sn_ctx* ctx, const sn_plane_job* jobs, int njobs) { SN_NOTHROW(ctx, process_planes_impl(ctx, jobs, njobs)) }

}  // extern "C"
