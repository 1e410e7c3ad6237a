// AviSynth filter class of the B200 SangNom2 plugin.
//
// Keeps the reference's constructor signature and script surface
// (/root/reference/src/SangNom2.h:40-67, SangNom2.cpp:275-330) but owns no scratch pool and no
// CPU kernels: GetFrame is a thin host layer that batches frames into libsangnom_cuda
// (include/sangnom_cuda.h). Uses only AviSynth+ API that exists with the same meaning in the real
// avisynth.h, so it builds against the SDK header as well as against the tests' stand-in tests/fakehost_src/avs_stub/avisynth.h.
#pragma once

#include <cstdint>
#include <deque>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "avisynth.h"
#include "sangnom_cuda.h"

class SangNom2 : public GenericVideoFilter {
    int order_;
    bool dh_;
    float aaf_[3];            // per-plane threshold in sample units, untruncated (reference :280-282)
    bool process_plane_[3];
    bool has_at_least_v8_;
    int sample_bytes_;
    int plane_count_;         // min(NumComponents, 3): planes the reference touches (:347)
    bool has_alpha_;

    sn_ctx* ctx_ = nullptr;
    int batch_frames_;        // frames fetched and processed per cache miss on sequential access
    int last_request_ = -2;
    int prefetch_depth_ = 2;  // batches submitted ahead of the one the host is consuming (0 = none)
    std::map<int, PVideoFrame> ready_;   // finished frames not yet (or recently) served
    std::mutex mu_;

    // Batches that have been submitted to the device but not waited for yet, oldest first: while the host consumes the
    // frames of batch k, batch k+1 is running and batch k+2 uploading (sangnom_cuda_submit / _wait) - with one batch
    // ahead only, the device pipeline drains between batches whenever the consumer is faster than a batch's latency.
    struct Pending {
        int first = 0, count = 0;
        sn_ticket ticket = 0;
        std::vector<PVideoFrame> srcs, dsts;   // keep the frame buffers alive until the batch is waited for
    };
    std::deque<Pending> pending_;

    // Frame buffers the host keeps recycling (AviSynth+'s frame registry hands the same VideoFrameBuffers out again
    // and again) are pinned for DMA the second time they are seen, so that their planes travel without the staging
    // copies of pageable memory; least recently used ones are unpinned when the budget is exceeded - only buffers that
    // have not been handed to the filter for many batches, so never one a batch in flight still uses.
    // SANGNOM_B200_PIN_MB=0 switches it off (see INTEGRATION.md for when to do that).
    struct PinEntry { size_t bytes = 0; int seen = 0; uint64_t last_use = 0; bool pinned = false; };
    std::unordered_map<const void*, PinEntry> pins_;
    size_t pinned_bytes_ = 0, pin_budget_ = 0, pin_min_bytes_ = 0;
    uint64_t pin_clock_ = 0;
    void note_frame_buffer(const PVideoFrame& f);

    int field_offset(int n);
    void start_batch(int first, int count, IScriptEnvironment* env);
    void finish_oldest_batch(int wanted, IScriptEnvironment* env);

public:
    SangNom2(PClip _child, int order, int aa, int aac, int threads, bool dh, bool luma, bool chroma, int opt, IScriptEnvironment* env);
    ~SangNom2() override;
    PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override;
    // One instance serves all host threads: GetFrame serialises on an internal lock and the
    // device work is batched, so cloning one GPU context per thread (MT_MULTI_INSTANCE, what the
    // reference answers at SangNom2.h:63-66 because of its per-instance scratch pool) would only
    // multiply device memory.
    int __stdcall SetCacheHints(int cachehints, int frame_range) override
    {
        (void)frame_range;
        return cachehints == CACHE_GET_MTMODE ? MT_NICE_FILTER : 0;
    }
};
