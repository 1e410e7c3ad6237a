// Worker pool for the host-side row copies of libsangnom_cuda (pure C++, no CUDA): used by sangnom_api.cu and, on its
// own, by tests/test_copy_pool.py through tests/emul/.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace sn_host {

// Host-side row copies (packing pageable frames into pinned staging, unpacking finished planes, whole-plane copies of
// planes that are not interpolated - the reference's BitBlt/memcpy, SangNom2.cpp:361-377) are memory-bound and a single
// thread moves only a few GB/s, far less than the PCIe link behind it. They are collected per chunk and run by a
// small pool of worker threads; the calling thread takes part.
struct RowCopy { char* dst; const char* src; ptrdiff_t dst_pitch, src_pitch; size_t row_bytes; int rows; };

class CopyPool {
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable wake_, done_;
    const std::vector<RowCopy>* jobs_ = nullptr;
    std::atomic<size_t> next_{ 0 };
    size_t total_ = 0;
    int active_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
    static constexpr int kRowsPerPiece = 64;

    // One row. Rows of a frame are written once and not read again by this CPU soon (the staging buffer is read by the
    // DMA engine, a finished frame by whoever consumes it later), so long rows go out with streaming stores: no
    // read-for-ownership of the destination lines, a third less memory traffic on a copy that is bandwidth-bound.
    static void copy_row(char* dst, const char* src, size_t n)
    {
#if defined(__SSE2__)
        if (n >= 512) {
            const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
            std::memcpy(dst, src, head);
            dst += head; src += head; n -= head;
            size_t i = 0;
            for (; i + 64 <= n; i += 64) {
                const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
                const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
                const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
                const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
            }
            std::memcpy(dst + i, src + i, n - i);
            return;
        }
#endif
        std::memcpy(dst, src, n);
    }
    static void copy_piece(const RowCopy& c, int r0, int r1)
    {
        for (int y = r0; y < r1; ++y) copy_row(c.dst + (ptrdiff_t)y * c.dst_pitch, c.src + (ptrdiff_t)y * c.src_pitch, c.row_bytes);
#if defined(__SSE2__)
        _mm_sfence();                           // streaming stores are visible before the piece counts as done
#endif
    }
    // pieces are numbered across all jobs: job j contributes ceil(rows / kRowsPerPiece) of them
    void drain(const std::vector<RowCopy>& jobs, const std::vector<size_t>& first_piece)
    {
        for (;;) {
            const size_t piece = next_.fetch_add(1, std::memory_order_relaxed);
            if (piece >= total_) return;
            const size_t j = (size_t)(std::upper_bound(first_piece.begin(), first_piece.end(), piece) - first_piece.begin()) - 1;
            const int r0 = (int)(piece - first_piece[j]) * kRowsPerPiece;
            copy_piece(jobs[j], r0, std::min(jobs[j].rows, r0 + kRowsPerPiece));
        }
    }
    std::vector<size_t> first_piece_;
    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                wake_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            drain(*jobs_, first_piece_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--active_ == 0) done_.notify_all();
            }
        }
    }
public:
    explicit CopyPool(int threads)
    {
        for (int i = 0; i < threads; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        wake_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // Runs every copy of `jobs`; returns when all are done. One caller at a time (the context lock is held).
    void run(const std::vector<RowCopy>& jobs)
    {
        if (jobs.empty()) return;
        first_piece_.assign(jobs.size() + 1, 0);
        size_t bytes = 0;
        for (size_t j = 0; j < jobs.size(); ++j) {
            first_piece_[j + 1] = first_piece_[j] + (size_t)(jobs[j].rows + kRowsPerPiece - 1) / kRowsPerPiece;
            bytes += jobs[j].row_bytes * (size_t)jobs[j].rows;
        }
        total_ = first_piece_.back();
        first_piece_.pop_back();
        next_.store(0, std::memory_order_relaxed);
        if (workers_.empty() || bytes < (size_t)1 << 20) { drain(jobs, first_piece_); return; }     // small: not worth waking anyone
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_ = &jobs;
            active_ = (int)workers_.size();
            ++generation_;
        }
        wake_.notify_all();
        drain(jobs, first_piece_);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return active_ == 0; });
    }
};

}  // namespace sn_host
