// Internal structures of libsangnom_cuda shared by the C-ABI front end (sangnom_api.cu) and the per-device host
// pipeline (sangnom_pipeline.cu). Nothing here is part of the public interface (include/sangnom_cuda.h).
//
// Shape of the host path:
//   caller thread   sangnom_cuda_submit: validate + plan the job list (frames, passes, cost-state hand-over), cut it
//                   into chunks of consecutive frames and deal the chunks round-robin to the pipelines
//   pipeline thread one per device: for every chunk  pack (host copies on the pipeline's own copy pool) -> H2D of the
//                   kept rows -> row-sweep kernels -> D2H of the interpolated rows, four chunks in flight; when a chunk's
//                   download has landed, unpack (pageable destinations) and tick the batch
//   caller thread   sangnom_cuda_wait: sleeps until every chunk of the batch (and of all earlier batches) has ticked
#pragma once
#include "sangnom_cuda.h"
#include "sangnom_kernels.h"
#include "sangnom_plan.h"
#include "host_copy_pool.h"

#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace sn_host {

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
        cudaError_t e = cudaHostAlloc(&p, need, cudaHostAllocPortable);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
};

// A processed plane of one frame, after validation.
struct Pass {
    const sn_plane_job* job = nullptr;
    int W = 0, H = 0, n = 0;     // samples, rows, kept rows
    int R = 0;                   // pool rows to sweep
    int cone = 0, export_cone = 0;   // dependency-cone bounds (sangnom_plan.h)
    sn::CostState in{}, out{};
    // host path: how the kept rows get up and the interpolated rows get down
    //   STAGED  pageable host memory: the pipeline's copy pool packs / scatters rows through the slot's pinned staging,
    //           one contiguous DMA transfer per run of planes
    //   LINEAR  pinned, the kept rows are contiguous (separated-field input with pitch == row): contiguous DMA straight
    //           from user memory, merged with its neighbours inside one pinned allocation
    //   PITCHED pinned, anything else: one 2-D DMA transfer (every other row of the frame)
    enum Xfer { STAGED, LINEAR, PITCHED };
    Xfer up = STAGED, down = STAGED;
    size_t src_off = 0, src_pitch = 0;       // device copy of the kept rows (bytes): n rows
    size_t out_off = 0, out_pitch = 0;       // device rows of the interpolated field: n + 1 rows (see Pipeline::start_chunk)
    size_t stage_in_off = 0, stage_out_off = 0;
    int src_pinned = 0, dst_pinned = 0;      // 0 = pageable, else 1 + pinned allocation index (PinnedLookup)
};

struct FramePlan {
    int key = 0;
    std::vector<Pass> passes;                // processed planes in plane order (at most 3)
    std::vector<const sn_plane_job*> copies; // planes that are only copied (disabled planes, alpha)
    size_t state_bytes = 0, state_off = 0;   // cost-state scratch of this frame
};

// One submitted job list. The jobs are copied so that the plans may outlive the caller's array.
struct Batch {
    uint64_t ticket = 0;
    std::vector<sn_plane_job> jobs;
    std::vector<FramePlan> frames;
    std::atomic<int> chunks_left{ 0 };
    std::mutex mu;                           // guards status / error
    int status = SN_OK;
    std::string error;
    void fail(int code, const std::string& msg)
    {
        std::lock_guard<std::mutex> lk(mu);
        if (status == SN_OK) { status = code; error = msg; }
    }
};

struct Chunk { Batch* batch = nullptr; size_t first = 0, last = 0; };      // frames [first, last) of the batch

constexpr int kSlots = 4;

// One pipeline slot: a chunk of frames resident on the device.
struct Slot {
    DevBuf planes, state, tasks;
    PinnedBuf tasks_host, stage_in, stage_out;
    cudaStream_t compute = nullptr, h2d = nullptr, d2h = nullptr;
    cudaEvent_t h2d_done = nullptr, kernels_done = nullptr, d2h_done = nullptr;
    cudaEvent_t t_h2d0 = nullptr, t_k0 = nullptr, t_d2h0 = nullptr;          // SANGNOM_TRACE only
    bool busy = false;
    Chunk chunk;
};

}  // namespace sn_host

struct sn_ctx;

namespace sn_host {

// The host pipeline of one device.
class Pipeline {
public:
    Pipeline(sn_ctx* ctx, int device, int copy_threads);
    ~Pipeline();
    cudaError_t init();                          // streams, events (on the calling thread, before start())
    void start();                                // spawn the worker
    void push(const Chunk& c);
    void stop_and_join();                        // finishes everything queued first
    int device() const { return device_; }

private:
    void run();
    int start_chunk(Slot& s, const Chunk& c, std::string& err);
    int finish_slot(Slot& s, std::string& err);
    void chunk_done(const Chunk& c, int status, const std::string& err);

    sn_ctx* ctx_;
    int device_;
    cudaStream_t h2d_ = nullptr, d2h_ = nullptr;
    cudaEvent_t trace_base_ = nullptr;
    Slot slots_[kSlots];
    int next_slot_ = 0;                          // slots are used round-robin, so this is also the oldest one in flight
    CopyPool pool_;
    std::thread worker_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Chunk> queue_;
    bool stop_ = false;
};

// Which pinned allocation a host pointer lies in (0 = pageable). DMA transfers may only be merged inside one
// allocation, so pinned-ness and the allocation's extent are looked up together, one driver query per allocation.
class PinnedLookup {
public:
    PinnedLookup();
    int find(const void* p);
private:
    struct Range { uintptr_t lo = 0, hi = 0; };
    void* get_ = nullptr;
    std::vector<Range> known_;
};

}  // namespace sn_host

constexpr int kTaskRing = 8;

struct sn_ctx {
    sn_config cfg{};
    int sample_bytes = 1;
    int S = 0, Hb = 0;
    int frames_in_flight = 0;
    bool persistent = false;                 // SN_FLAG_PERSISTENT_POOL
    bool saturate = false;                   // SN_FLAG_SATURATE
    std::vector<int> devices;                // the device of each pipeline; devices[0] also serves the device entry
    std::vector<std::unique_ptr<sn_host::Pipeline>> pipelines;
    size_t next_pipeline = 0;                // chunks are dealt round-robin

    // batches in flight, oldest first; guarded by mu. done_cv is signalled whenever a batch's last chunk finishes.
    std::mutex mu;
    std::condition_variable done_cv;
    uint64_t last_ticket = 0;
    std::deque<std::unique_ptr<sn_host::Batch>> batches;

    // device-entry resources (devices[0]); guarded by dev_mu
    std::mutex dev_mu;
    cudaStream_t own_compute = nullptr;
    sn_host::DevBuf dev_state;
    sn_host::DevBuf dev_tasks[kTaskRing];
    sn_host::PinnedBuf dev_tasks_host[kTaskRing];
    cudaEvent_t dev_task_free[kTaskRing] = {};
    int dev_ring_pos = 0;
    // persistent-pool mode: the pool state between frames, ping-pong; guarded by carry_mu (frames are chained in the
    // order their tasks are built)
    std::mutex carry_mu;
    sn_host::DevBuf carry[2];
    int carry_pos = 0;                       // carry[carry_pos] holds the state the next frame starts from

    // host ranges pinned through sangnom_cuda_host_pin; guarded by mu
    struct Pinned { void* base; size_t bytes; };
    std::vector<Pinned> pinned;

    std::mutex stats_mu;
    sn_stats stats{};
    std::mutex err_mu;
    std::string error;

    int fail(int code, const char* fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        std::lock_guard<std::mutex> lk(err_mu);
        error = buf;
        return code;
    }
    int cuda_fail(cudaError_t e, const char* what) { return fail(SN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e)); }
};

namespace sn_host {

// Shared by both entries (sangnom_api.cu).
sn::PlaneTask make_task(const sn_ctx* ctx, const Pass& p, void* plane, size_t pitch_bytes, const void* kept0, size_t kept_step_bytes, int copy_kept);
void place_state(sn_ctx* ctx, FramePlan& f, char* base);
void add_task(const sn_ctx* ctx, std::vector<std::vector<sn::PlaneTask>>& launches, size_t q, const sn::PlaneTask& t);
// Copies the task arrays to `host_tasks` and queues their upload; then one launch per entry of by_pass.
cudaError_t upload_tasks(const std::vector<std::vector<sn::PlaneTask>>& by_pass, sn::PlaneTask* host_tasks, sn::PlaneTask* dev_tasks, cudaStream_t stream);
cudaError_t launch_passes(sn_ctx* ctx, const std::vector<std::vector<sn::PlaneTask>>& by_pass, sn::PlaneTask* dev_tasks, cudaStream_t stream);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace sn_host
