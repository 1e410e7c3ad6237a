// Host pipeline of one device (see sangnom_ctx.h for the shape of the host path).
//
// What crosses PCIe per processed plane is the minimum the reference's seam implies (/root/reference/src/
// SangNom2.cpp:361-393): UP the kept field (the rows the BitBlt at :361-377 copies, W*H/2 samples), DOWN the
// interpolated rows (W*(H/2-1) samples). The kept rows and the border row of the destination frame (:380-391) never
// leave the host: they are copied source -> destination by the pipeline's copy pool while the DMA engines work.
// The device holds, per plane, the uploaded kept rows and a packed block of interpolated rows; the kernel is pointed
// at that block through a plane pointer / pitch pair under which picture row offset+1+2j is packed row j.
#include "sangnom_ctx.h"

#include <cuda.h>      // driver API types only; the entry point is looked up at run time (no libcuda link dependency)

#include <algorithm>
#include <chrono>
#include <cstring>

namespace sn_host {

// ---- PinnedLookup -------------------------------------------------------------------------------------------------
namespace {
using GetAttr = CUresult (*)(void*, CUpointer_attribute, CUdeviceptr);
GetAttr driver_entry()          // looked up once per process
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
}  // namespace

PinnedLookup::PinnedLookup() : get_(reinterpret_cast<void*>(driver_entry())) {}

int PinnedLookup::find(const void* p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    for (size_t i = 0; i < known_.size(); ++i)
        if (a >= known_[i].lo && a < known_[i].hi) return (int)i + 1;
    Range r;
    if (get_) {
        const GetAttr get = reinterpret_cast<GetAttr>(get_);
        CUmemorytype type{};
        CUdeviceptr base = 0;
        size_t size = 0;
        if (get(&type, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)a) != CUDA_SUCCESS || type != CU_MEMORYTYPE_HOST) return 0;
        if (get(&base, CU_POINTER_ATTRIBUTE_RANGE_START_ADDR, (CUdeviceptr)a) != CUDA_SUCCESS ||
            get(&size, CU_POINTER_ATTRIBUTE_RANGE_SIZE, (CUdeviceptr)a) != CUDA_SUCCESS || size == 0) {
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

// ---- DMA runs -------------------------------------------------------------------------------------------------------
namespace {

// A run of bytes that is contiguous both in (pinned) host memory and in the device slot: one DMA transfer.
struct Segment { char* host; char* dev; size_t bytes; int alloc; };

// alloc: which pinned allocation `host` lies in (runs are merged only inside one allocation: a copy that spans two
// cudaHostAlloc blocks fails even when they are neighbours in the address space)
void add_segment(std::vector<Segment>& v, void* host, void* dev, size_t bytes, int alloc)
{
    if (bytes == 0) return;
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

const bool g_trace = getenv("SANGNOM_TRACE") != nullptr;
// SANGNOM_B200_COPY_STREAMS=shared: all chunks of a pipeline upload on one stream and download on another (one DMA
// engine per direction). Default: every slot has its own pair, so that the 2-D transfers of consecutive chunks can
// run on different copy engines at the same time.
const bool g_shared_copy_streams = [] { const char* v = getenv("SANGNOM_B200_COPY_STREAMS"); return v && std::strcmp(v, "shared") == 0; }();

// Planes of consecutive frames that a batching host layer staged back to back in one pinned arena sit at a constant
// stride: the same rows of all of them are ONE 3-D transfer (width = row bytes, height = rows, depth = frames) instead
// of one 2-D transfer per plane. On the copy engines of this box that is the difference between 26 and 41 GB/s per
// direction with both directions busy (tools/dma_probe.cu): a 2-D copy of a few hundred rows is dominated by its
// set-up. A run is formed per (pass index, geometry, field offset): with alternating field parity the even and the
// odd frames of a clip form two interleaved runs.
struct Run {
    std::vector<Pass*> passes;
    const char* first = nullptr;     // host address of the first row of the first plane
    ptrdiff_t stride = 0;            // host bytes from one plane of the run to the next (0: single plane so far)
    size_t row = 0, step = 0;        // row bytes, host bytes between consecutive rows
    int rows = 0, q = 0, offset = 0, alloc = 0;
};

// Append plane `p` (rows start at `host`) to the open run with the same key, or open a new one.
void add_to_runs(std::vector<Run>& runs, Pass* p, int q, const char* host, size_t row, size_t step, int rows, int alloc)
{
    for (auto it = runs.rbegin(); it != runs.rend(); ++it) {
        Run& r = *it;
        if (r.q != q || r.offset != p->job->offset || r.row != row || r.step != step || r.rows != rows || r.alloc != alloc) continue;
        const char* last = r.first + (ptrdiff_t)(r.passes.size() - 1) * r.stride;
        const ptrdiff_t d = host - last;
        if (r.passes.size() == 1) {
            // the slice pitch of a 3-D copy is a whole number of rows and must hold the rows copied
            if (d > 0 && (size_t)d % step == 0 && (size_t)d / step >= (size_t)rows) { r.stride = d; r.passes.push_back(p); return; }
        } else if (d == r.stride) { r.passes.push_back(p); return; }
        break;                                   // the newest run with this key does not continue: start another
    }
    Run r;
    r.passes.push_back(p); r.first = host; r.row = row; r.step = step; r.rows = rows; r.q = q; r.offset = p->job->offset; r.alloc = alloc;
    runs.push_back(r);
}

// The kept rows of a job in host memory: first row and the step between consecutive kept rows (bytes).
inline const char* kept_rows(const sn_plane_job& jb, ptrdiff_t& step)
{
    const bool field = jb.mode == SN_MODE_FIELD;
    step = jb.src_pitch * (field ? 2 : 1);
    return static_cast<const char*>(jb.src) + (field ? (ptrdiff_t)jb.offset * jb.src_pitch : 0);
}

}  // namespace

// ---- Pipeline -------------------------------------------------------------------------------------------------------
Pipeline::Pipeline(sn_ctx* ctx, int device, int copy_threads) : ctx_(ctx), device_(device), pool_(std::max(0, copy_threads - 1)) {}

Pipeline::~Pipeline()
{
    stop_and_join();
    cudaSetDevice(device_);
    for (Slot& s : slots_) {
        s.planes.release(); s.state.release(); s.tasks.release();
        s.tasks_host.release(); s.stage_in.release(); s.stage_out.release();
        if (s.compute) cudaStreamDestroy(s.compute);
        if (s.h2d && s.h2d != h2d_) cudaStreamDestroy(s.h2d);
        if (s.d2h && s.d2h != d2h_) cudaStreamDestroy(s.d2h);
        for (cudaEvent_t e : { s.h2d_done, s.kernels_done, s.d2h_done, s.t_h2d0, s.t_k0, s.t_d2h0 })
            if (e) cudaEventDestroy(e);
    }
    if (trace_base_) cudaEventDestroy(trace_base_);
    if (h2d_) cudaStreamDestroy(h2d_);
    if (d2h_) cudaStreamDestroy(d2h_);
    cudaGetLastError();
}

cudaError_t Pipeline::init()
{
    cudaError_t e;
    if ((e = cudaSetDevice(device_)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&h2d_, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&d2h_, cudaStreamNonBlocking)) != cudaSuccess) return e;
    for (Slot& s : slots_) {
        if ((e = cudaStreamCreateWithFlags(&s.compute, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if (g_shared_copy_streams || ctx_->persistent) { s.h2d = h2d_; s.d2h = d2h_; }
        else {
            if ((e = cudaStreamCreateWithFlags(&s.h2d, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaStreamCreateWithFlags(&s.d2h, cudaStreamNonBlocking)) != cudaSuccess) return e;
        }
        for (cudaEvent_t* ev : { &s.h2d_done, &s.kernels_done, &s.d2h_done })
            if ((e = cudaEventCreateWithFlags(ev, g_trace ? cudaEventDefault : cudaEventDisableTiming)) != cudaSuccess) return e;
        if (g_trace)
            for (cudaEvent_t* ev : { &s.t_h2d0, &s.t_k0, &s.t_d2h0 })
                if ((e = cudaEventCreate(ev)) != cudaSuccess) return e;
    }
    if (g_trace) {
        if ((e = cudaEventCreate(&trace_base_)) != cudaSuccess) return e;
        cudaEventRecord(trace_base_, h2d_);
    }
    return cudaSuccess;
}

void Pipeline::start() { worker_ = std::thread([this] { run(); }); }

void Pipeline::push(const Chunk& c)
{
    { std::lock_guard<std::mutex> lk(mu_); queue_.push_back(c); }
    cv_.notify_one();
}

void Pipeline::stop_and_join()
{
    if (!worker_.joinable()) return;
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_one();
    worker_.join();
}

void Pipeline::chunk_done(const Chunk& c, int status, const std::string& err)
{
    if (status != SN_OK) c.batch->fail(status, err);
    else {
        std::lock_guard<std::mutex> lk(ctx_->stats_mu);
        ctx_->stats.frames += c.last - c.first;
    }
    if (c.batch->chunks_left.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> lk(ctx_->mu);          // the waiter checks chunks_left under this lock
        ctx_->done_cv.notify_all();
    }
}

void Pipeline::run()
{
    cudaSetDevice(device_);
    for (;;) {
        Chunk c;
        bool have = false;
        bool any_busy = false;
        for (const Slot& s : slots_) any_busy = any_busy || s.busy;
        {
            std::unique_lock<std::mutex> lk(mu_);
            if (!any_busy) cv_.wait(lk, [&] { return stop_ || !queue_.empty(); });
            if (!queue_.empty()) { c = queue_.front(); queue_.pop_front(); have = true; }
            else if (stop_ && !any_busy) return;
        }
        std::string err;
        if (have) {
            Slot& s = slots_[next_slot_];
            next_slot_ = (next_slot_ + 1) % kSlots;
            if (s.busy) {                                   // the oldest chunk in flight: its buffers are needed
                const Chunk old = s.chunk;
                const int rc = finish_slot(s, err);
                chunk_done(old, rc, err);
                err.clear();
            }
            const int rc = start_chunk(s, c, err);
            if (rc != SN_OK) chunk_done(c, rc, err);
        } else {
            // nothing queued: bring the oldest chunk in flight home
            for (int k = 0; k < kSlots; ++k) {
                Slot& s = slots_[(next_slot_ + k) % kSlots];
                if (!s.busy) continue;
                const Chunk old = s.chunk;
                const int rc = finish_slot(s, err);
                chunk_done(old, rc, err);
                break;
            }
        }
    }
}

#define PL_CUDA(call, what)                                                                          \
    do {                                                                                             \
        const cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) { err = std::string(what) + ": " + cudaGetErrorString(e__); status = SN_ERR_CUDA; break; } \
    } while (0)

int Pipeline::start_chunk(Slot& s, const Chunk& c, std::string& err)
{
    sn_ctx* const ctx = ctx_;
    const int sb = ctx->sample_bytes;
    std::vector<FramePlan>& frames = c.batch->frames;
    int status = SN_OK;

    // ---- placement inside the slot: one region for the uploaded kept rows and one for the packed interpolated rows,
    // each filled in job order at 16 / 32-byte granules, so that planes which are neighbours inside one pinned host
    // allocation (or in the slot's own staging) are neighbours on the device and their DMA transfers merge into one
    size_t src_total = 0, out_total = 0, state_bytes = 0, in_bytes = 0, out_bytes = 0, ntasks = 0;
    std::vector<Run> up_runs, down_runs;             // pinned planes that travel by 2-D / 3-D DMA
    for (size_t k = c.first; k < c.last; ++k) {
        FramePlan& f = frames[k];
        f.state_off = state_bytes;
        state_bytes += align_up(f.state_bytes, 256);
        for (size_t q = 0; q < f.passes.size(); ++q) {
            Pass& p = f.passes[q];
            const sn_plane_job& jb = *p.job;
            const size_t row = (size_t)p.W * sb;
            // 32-byte rows: half a row (the pitch the kernel is given, see below) stays a multiple of its widest store
            p.out_pitch = align_up(row, 32);
            if (!p.src_pinned) { p.up = Pass::STAGED; p.src_pitch = align_up(row, 16); p.stage_in_off = in_bytes; in_bytes += p.src_pitch * p.n; }
            else if (jb.mode != SN_MODE_FIELD && (size_t)jb.src_pitch == row && row % 16 == 0) { p.up = Pass::LINEAR; p.src_pitch = row; }
            else {
                p.up = Pass::PITCHED; p.src_pitch = align_up(row, 16);
                ptrdiff_t step;
                const char* kept = kept_rows(jb, step);
                add_to_runs(up_runs, &p, (int)q, kept, row, (size_t)step, p.n, p.src_pinned);
            }
            if (!p.dst_pinned) { p.down = Pass::STAGED; p.stage_out_off = out_bytes; out_bytes += p.out_pitch * (size_t)(p.n - 1); }
            else {
                p.down = Pass::PITCHED;
                if (p.n > 1) add_to_runs(down_runs, &p, (int)q, static_cast<const char*>(jb.dst) + (ptrdiff_t)(jb.offset + 1) * jb.dst_pitch, row, 2 * (size_t)jb.dst_pitch, p.n - 1, p.dst_pinned);
            }
            // staged and linear planes keep job order (that is what lets their transfers merge); run members are placed below
            if (p.up != Pass::PITCHED) { p.src_off = src_total; src_total += p.src_pitch * p.n; }
            if (p.down != Pass::PITCHED || p.n < 2) { p.out_off = out_total; out_total += p.out_pitch * (size_t)(p.n - 1); }
            ++ntasks;
        }
    }
    // the planes of a run back to back on the device: the device side of the 3-D transfer has a slice of exactly `rows` rows
    for (Run& r : up_runs)
        for (Pass* p : r.passes) { p->src_off = src_total; src_total += p->src_pitch * p->n; }
    for (Run& r : down_runs)
        for (Pass* p : r.passes) { p->out_off = out_total; out_total += p->out_pitch * (size_t)(p->n - 1); }
    // the kernel's border-row store is switched off (no_border), but rows are addressed relative to a pointer two
    // packed half-rows before the block: keep a margin in front of it inside the allocation
    const size_t out_base = align_up(src_total, 256) + 256;
    const size_t plane_bytes = out_base + out_total + 256;

    do {
        cudaError_t e;
        if ((e = s.planes.ensure(plane_bytes)) != cudaSuccess || (e = s.state.ensure(state_bytes)) != cudaSuccess ||
            (e = s.tasks.ensure(ntasks * sizeof(sn::PlaneTask))) != cudaSuccess ||
            (e = s.tasks_host.ensure(ntasks * sizeof(sn::PlaneTask))) != cudaSuccess ||
            (e = s.stage_in.ensure(in_bytes)) != cudaSuccess || (e = s.stage_out.ensure(out_bytes)) != cudaSuccess) {
            err = std::string("slot allocation: ") + cudaGetErrorString(e);
            status = (e == cudaErrorMemoryAllocation) ? SN_ERR_NOMEM : SN_ERR_CUDA;
            cudaGetLastError();
            break;
        }
        char* const dplanes = static_cast<char*>(s.planes.p);

        // ---- host-side copies of the chunk, all at once on the copy pool: the kept rows and the border row of every
        // destination plane (reference :361-391), the planes that are only copied (disabled planes, alpha, :369-374),
        // and the kept rows of pageable sources into the pinned staging buffer
        uint64_t host_bytes = 0;
        {
            std::vector<RowCopy> copies;
            for (size_t k = c.first; k < c.last; ++k) {
                FramePlan& f = frames[k];
                for (const sn_plane_job* cj : f.copies) {
                    if (cj->src == cj->dst) continue;
                    copies.push_back(RowCopy{ static_cast<char*>(cj->dst), static_cast<const char*>(cj->src), cj->dst_pitch, cj->src_pitch, (size_t)cj->width * sb, cj->dst_height });
                    host_bytes += (uint64_t)cj->width * sb * cj->dst_height;
                }
                for (Pass& p : f.passes) {
                    const sn_plane_job& jb = *p.job;
                    const size_t row = (size_t)p.W * sb;
                    ptrdiff_t step;
                    const char* kept = kept_rows(jb, step);
                    char* const dst = static_cast<char*>(jb.dst);
                    char* const dkept = dst + (ptrdiff_t)jb.offset * jb.dst_pitch;
                    if (!(kept == dkept && step == 2 * jb.dst_pitch)) {          // not already in place in the destination frame
                        copies.push_back(RowCopy{ dkept, kept, 2 * jb.dst_pitch, step, row, p.n });
                        host_bytes += (uint64_t)row * p.n;
                    }
                    // the row without a neighbour pair: a copy of the nearest kept row (:380-391)
                    if (jb.offset == 0) copies.push_back(RowCopy{ dst + (ptrdiff_t)(p.H - 1) * jb.dst_pitch, kept + (ptrdiff_t)(p.n - 1) * step, 0, 0, row, 1 });
                    else copies.push_back(RowCopy{ dst, kept, 0, 0, row, 1 });
                    if (p.up == Pass::STAGED) {
                        copies.push_back(RowCopy{ static_cast<char*>(s.stage_in.p) + p.stage_in_off, kept, (ptrdiff_t)p.src_pitch, step, row, p.n });
                        host_bytes += (uint64_t)row * p.n;
                    }
                }
            }
            pool_.run(copies);
        }

        // ---- upload of the kept rows ----
        uint64_t h2d_bytes = 0, d2h_bytes = 0;
        if (g_trace) cudaEventRecord(s.t_h2d0, s.h2d);
        std::vector<std::vector<sn::PlaneTask>> by_pass;
        std::vector<Segment> up_segs, down_segs;
        for (size_t k = c.first; k < c.last && status == SN_OK; ++k) {
            FramePlan& f = frames[k];
            place_state(ctx, f, static_cast<char*>(s.state.p) + f.state_off);
            for (size_t q = 0; q < f.passes.size(); ++q) {
                Pass& p = f.passes[q];
                const sn_plane_job& jb = *p.job;
                const size_t row = (size_t)p.W * sb;
                ptrdiff_t step;
                const char* kept = kept_rows(jb, step);
                char* const dsrc = dplanes + p.src_off;
                if (p.up == Pass::STAGED) add_segment(up_segs, static_cast<char*>(s.stage_in.p) + p.stage_in_off, dsrc, p.src_pitch * p.n, -1);
                else if (p.up == Pass::LINEAR) add_segment(up_segs, const_cast<char*>(kept), dsrc, row * p.n, p.src_pinned);
                h2d_bytes += p.up == Pass::PITCHED ? row * p.n : p.src_pitch * p.n;           // (pitched planes travel with their run, below)
                // Picture row offset+1+2j of the plane is packed row j of the block at out_off: hand the kernel a pitch of
                // half a packed row and a plane pointer (offset+1) half-rows before the block.
                char* const dout = dplanes + out_base + p.out_off;
                const size_t half = p.out_pitch / 2;
                add_task(ctx, by_pass, q, make_task(ctx, p, dout - (size_t)(jb.offset + 1) * half, half, dsrc, p.src_pitch, /*copy_kept=*/0));
            }
        }
        if (status != SN_OK) break;
        PL_CUDA(flush_segments(up_segs, cudaMemcpyHostToDevice, s.h2d), "H2D copy");
        for (const Run& r : up_runs) {
            const Pass& p0 = *r.passes.front();
            char* const dev = dplanes + p0.src_off;
            if (r.passes.size() == 1) {
                PL_CUDA(cudaMemcpy2DAsync(dev, p0.src_pitch, r.first, r.step, r.row, (size_t)r.rows, cudaMemcpyHostToDevice, s.h2d), "H2D copy");
            } else {
                cudaMemcpy3DParms q3{};
                q3.srcPtr = make_cudaPitchedPtr(const_cast<char*>(r.first), r.step, r.row, (size_t)r.stride / r.step);
                q3.dstPtr = make_cudaPitchedPtr(dev, p0.src_pitch, r.row, (size_t)r.rows);
                q3.extent = make_cudaExtent(r.row, (size_t)r.rows, r.passes.size());
                q3.kind = cudaMemcpyHostToDevice;
                PL_CUDA(cudaMemcpy3DAsync(&q3, s.h2d), "H2D copy");
            }
        }
        if (status != SN_OK) break;
        // The task array rides the upload stream too: a small copy on the compute stream would queue on the
        // same DMA engine behind the NEXT chunks' bulk uploads and hold this chunk's kernels back.
        PL_CUDA(upload_tasks(by_pass, static_cast<sn::PlaneTask*>(s.tasks_host.p), static_cast<sn::PlaneTask*>(s.tasks.p), s.h2d), "task upload");
        // persistent pool: every chunk's kernels on ONE stream, so that frames run in submission order across chunks
        const cudaStream_t compute = ctx->persistent ? slots_[0].compute : s.compute;
        PL_CUDA(cudaEventRecord(s.h2d_done, s.h2d), "event");
        PL_CUDA(cudaStreamWaitEvent(compute, s.h2d_done, 0), "event");

        // ---- kernels ----
        if (g_trace) cudaEventRecord(s.t_k0, compute);
        PL_CUDA(launch_passes(ctx, by_pass, static_cast<sn::PlaneTask*>(s.tasks.p), compute), "kernel launch");
        PL_CUDA(cudaEventRecord(s.kernels_done, compute), "event");
        PL_CUDA(cudaStreamWaitEvent(s.d2h, s.kernels_done, 0), "event");

        // ---- download of the interpolated rows ----
        if (g_trace) cudaEventRecord(s.t_d2h0, s.d2h);
        for (size_t k = c.first; k < c.last; ++k)
            for (Pass& p : frames[k].passes) {
                if (p.n < 2) continue;
                if (p.down == Pass::STAGED) {
                    add_segment(down_segs, static_cast<char*>(s.stage_out.p) + p.stage_out_off, dplanes + out_base + p.out_off, p.out_pitch * (size_t)(p.n - 1), -1);
                    d2h_bytes += p.out_pitch * (size_t)(p.n - 1);
                } else d2h_bytes += (size_t)p.W * sb * (size_t)(p.n - 1);
            }
        PL_CUDA(flush_segments(down_segs, cudaMemcpyDeviceToHost, s.d2h), "D2H copy");
        for (const Run& r : down_runs) {
            const Pass& p0 = *r.passes.front();
            char* const dev = dplanes + out_base + p0.out_off;
            if (r.passes.size() == 1) {
                PL_CUDA(cudaMemcpy2DAsync(const_cast<char*>(r.first), r.step, dev, p0.out_pitch, r.row, (size_t)r.rows, cudaMemcpyDeviceToHost, s.d2h), "D2H copy");
            } else {
                cudaMemcpy3DParms q3{};
                q3.srcPtr = make_cudaPitchedPtr(dev, p0.out_pitch, r.row, (size_t)r.rows);
                q3.dstPtr = make_cudaPitchedPtr(const_cast<char*>(r.first), r.step, r.row, (size_t)r.stride / r.step);
                q3.extent = make_cudaExtent(r.row, (size_t)r.rows, r.passes.size());
                q3.kind = cudaMemcpyDeviceToHost;
                PL_CUDA(cudaMemcpy3DAsync(&q3, s.d2h), "D2H copy");
            }
        }
        if (status != SN_OK) break;
        PL_CUDA(cudaEventRecord(s.d2h_done, s.d2h), "event");
        s.busy = true;
        s.chunk = c;
        {
            std::lock_guard<std::mutex> lk(ctx->stats_mu);
            ctx->stats.h2d_bytes += h2d_bytes;
            ctx->stats.d2h_bytes += d2h_bytes;
            ctx->stats.host_copy_bytes += host_bytes;
        }
    } while (0);

    if (status != SN_OK) {
        // leave nothing of this chunk in flight that reads or writes user memory
        cudaStreamSynchronize(s.h2d);
        cudaStreamSynchronize(ctx->persistent ? slots_[0].compute : s.compute);
        cudaStreamSynchronize(s.d2h);
        cudaGetLastError();
    }
    return status;
}

// Wait for the chunk's download and scatter the interpolated rows of pageable destinations out of the staging buffer.
int Pipeline::finish_slot(Slot& s, std::string& err)
{
    if (!s.busy) return SN_OK;
    s.busy = false;
    const cudaError_t e = cudaEventSynchronize(s.d2h_done);
    if (e != cudaSuccess) { err = std::string("waiting for the download: ") + cudaGetErrorString(e); cudaGetLastError(); return SN_ERR_CUDA; }
    if (g_trace) {
        float t[6] = {};
        cudaEventElapsedTime(&t[0], trace_base_, s.t_h2d0); cudaEventElapsedTime(&t[1], trace_base_, s.h2d_done);
        cudaEventElapsedTime(&t[2], trace_base_, s.t_k0); cudaEventElapsedTime(&t[3], trace_base_, s.kernels_done);
        cudaEventElapsedTime(&t[4], trace_base_, s.t_d2h0); cudaEventElapsedTime(&t[5], trace_base_, s.d2h_done);
        fprintf(stderr, "[sangnom] device %d chunk of %3zu frames: h2d %7.2f..%7.2f  kernels %7.2f..%7.2f  d2h %7.2f..%7.2f ms\n",
                device_, s.chunk.last - s.chunk.first, t[0], t[1], t[2], t[3], t[4], t[5]);
    }
    const int sb = ctx_->sample_bytes;
    std::vector<RowCopy> copies;
    uint64_t host_bytes = 0;
    std::vector<FramePlan>& frames = s.chunk.batch->frames;
    for (size_t k = s.chunk.first; k < s.chunk.last; ++k)
        for (Pass& p : frames[k].passes) {
            if (p.down != Pass::STAGED || p.n < 2) continue;
            const sn_plane_job& jb = *p.job;
            copies.push_back(RowCopy{ static_cast<char*>(jb.dst) + (ptrdiff_t)(jb.offset + 1) * jb.dst_pitch, static_cast<const char*>(s.stage_out.p) + p.stage_out_off,
                                      2 * jb.dst_pitch, (ptrdiff_t)p.out_pitch, (size_t)p.W * sb, p.n - 1 });
            host_bytes += (uint64_t)p.W * sb * (p.n - 1);
        }
    pool_.run(copies);
    if (host_bytes) {
        std::lock_guard<std::mutex> lk(ctx_->stats_mu);
        ctx_->stats.host_copy_bytes += host_bytes;
    }
    return SN_OK;
}

}  // namespace sn_host
