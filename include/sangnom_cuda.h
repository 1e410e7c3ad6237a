/*
 * sangnom_cuda.h - C ABI of libsangnom_cuda: SangNom2's edge-directed field interpolation on
 * NVIDIA B200 (sm_100a only; there is no CPU fallback - every entry point fails with
 * SN_ERR_CUDA when no usable device is present).
 *
 * This is the drop-in seam for the reference's per-plane hot path. Each entry point names the
 * reference interface it replaces (paths relative to /root/reference):
 *
 *   sangnom_cuda_create / _destroy
 *       replaces the scratch-pool half of SangNom2::SangNom2 (src/SangNom2.cpp:287-310: pool
 *       geometry bufferStride/bufferHeight from the OUTPUT luma size, pool + line allocation) and
 *       the implicit free in ~SangNom2 (src/SangNom2.h:50-51). The context owns the CUDA streams,
 *       pinned staging, device planes and the inter-plane cost-state scratch.
 *   sangnom_cuda_process_planes
 *       replaces, for a batch of planes, the body of the plane loop in SangNom2::GetFrame
 *       (src/SangNom2.cpp:348-394): the kept-field copy (:361-377), the un-interpolatable border
 *       row copy (:380-391) and the call through the `process` member pointer (:393), i.e.
 *       sangnom_c<T,IType> (:259-273) = prepareBuffers_c (:74-124) + 9 x processBuffers_c
 *       (:126-159) + finalizePlane_c (:161-257).
 *   sangnom_cuda_process_planes_device
 *       the same seam for planes already resident in device memory (no PCIe copies), in place
 *       exactly like `process(dstp, dstStride, w, h, offset, plane)` (:393).
 *   sangnom_cuda_threshold
 *       replaces the aaf[] scaling in the ctor (src/SangNom2.cpp:280-282).
 *
 * Numerics contract: results are bit-identical to the reference's opt=0 C++ path for 8..16-bit
 * integer samples and for fp32 (same operation order, no FMA contraction), under the rule
 * "each frame is processed as by a freshly constructed filter instance whose scratch pool is
 * zero-filled, planes in Y,U,V order" (DESIGN.md, section Parity contract). The coupling of the
 * U and V planes to the luma cost state that the reference's shared pool produces is reproduced
 * exactly; that is why jobs carry a frame key.
 */
#ifndef SANGNOM_CUDA_H
#define SANGNOM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SN_API __attribute__((visibility("default")))
#else
#define SN_API
#endif

#define SANGNOM_CUDA_ABI_VERSION 2

typedef struct sn_ctx sn_ctx;

enum sn_status {
    SN_OK = 0,
    SN_ERR_INVALID = 1,      /* bad argument / inconsistent job list */
    SN_ERR_CUDA = 2,         /* CUDA runtime error or no sm_100 device */
    SN_ERR_UNSUPPORTED = 3,  /* geometry outside what the kernels cover (see sn_limits) */
    SN_ERR_NOMEM = 4
};

enum sn_sample {             /* == bytes per sample == AviSynth ComponentSize() */
    SN_SAMPLE_U8 = 1,
    SN_SAMPLE_U16 = 2,       /* 10/12/14/16-bit ride in 16-bit containers */
    SN_SAMPLE_F32 = 4
};

enum sn_mode {
    SN_MODE_COPY = 0,        /* plane not processed: dst := src (disabled plane, alpha) (:369-374) */
    SN_MODE_FIELD = 1,       /* src is a full-height plane; rows offset,offset+2,.. are kept (:376) */
    SN_MODE_DH = 2,          /* src has dst_height/2 rows; they become rows offset,offset+2,.. (:361-366) */
    SN_MODE_INPLACE = 3      /* device entry only: dst already holds the kept field; src ignored (:393) */
};

/* sn_config.flags */
/* Persistent scratch pool: frames are processed strictly in submission order and each frame starts from the pool
 * state the previous frame left, exactly like ONE long-lived reference instance pulled sequentially (the reference
 * never clears its pool, SangNom2.cpp:303-310). Matters only where the reference is not frame-pure: luma width not a
 * multiple of 32 (the pad columns carry over) or luma=false with subsampled chroma. Frames become sequentially
 * dependent, so this mode runs one plane pass at a time: use it for bit-compatibility with a sequential reference
 * run, not for throughput. Default (0): every frame starts from a zero-filled pool (a fresh instance per frame). */
#define SN_FLAG_PERSISTENT_POOL 1
/* Arithmetic flavour of the reference's SSE2 path (opt=1, and what opt=-1 picks on an SSE2 CPU; SangNom2_SSE2.cpp):
 * the same three stages, but the 3-tap value and the blurred cost are narrowed with SATURATION (:449-517, :761, :807)
 * where the opt=0 C++ path wraps (SangNom2.cpp:63-64, :152). Integer formats only; fp32 is identical in both. Default
 * (0): opt=0 arithmetic, the parity contract. */
#define SN_FLAG_SATURATE 2

#define SN_DEVICE_ALL (-1)   /* sn_config.device: one pipeline on every sm_100 device the process can see */

typedef struct sn_config {
    int abi_version;         /* SANGNOM_CUDA_ABI_VERSION */
    int device;              /* CUDA device ordinal, or SN_DEVICE_ALL */
    int sample_type;         /* enum sn_sample */
    int pool_width;          /* OUTPUT luma width in samples  (vi.width)            -> S  = align32 */
    int pool_height;         /* OUTPUT luma height in rows    (vi.height after dh)  -> Hb = (h+1)>>1 */
    int max_frames_in_flight;/* frames resident on ONE device at once (0 = library default) */
    int flags;               /* SN_FLAG_* bits, 0 = defaults */
    /* Several devices behind one context (host entry). Frames are independent under the parity contract, so the
     * chunks of consecutive frames a batch is cut into are dealt round-robin to one pipeline per device - own
     * streams, staging and host thread each, no data exchanged between devices; a frame's planes always stay on one
     * device (U reads Y's cost state, V reads U's). This is what the reference gets from the host's MT model
     * (SangNom2.h:63-66, one filter instance per worker thread). Bit d set = use CUDA device d; 0 = `device` alone
     * (or all devices for SN_DEVICE_ALL). SN_FLAG_PERSISTENT_POOL chains the frames and therefore runs on the first
     * device only. The device entry always runs on the first device of the context. */
    unsigned long long device_mask;
    int copy_threads;        /* host threads (all pipelines together) for the row copies of the host entry; 0 = default */
} sn_config;

/* One plane of one frame. Pitches are in BYTES. */
typedef struct sn_plane_job {
    const void* src;
    ptrdiff_t src_pitch;
    void* dst;
    ptrdiff_t dst_pitch;
    int width;               /* samples per row */
    int dst_height;          /* rows of the dst plane (src has dst_height rows, or dst_height/2 for DH) */
    int offset;              /* 0: keep rows 0,2,4.. (top field)  1: keep rows 1,3,5.. */
    int mode;                /* enum sn_mode */
    float threshold;         /* sangnom_cuda_threshold(aa or aac, bits, sample_type) */
    int plane;               /* 0 Y, 1 U, 2 V, 3 A - processing order inside a frame is by this index */
    int frame;               /* caller's key: jobs with equal key belong to one frame and share one
                                (virtual) scratch pool, exactly like one GetFrame call */
} sn_plane_job;

typedef struct sn_limits {
    int max_pool_width[5];   /* indexed by sample bytes (1,2,4): widest pool the kernels accept */
    int sm_count;
    int compute_major, compute_minor;
} sn_limits;

typedef struct sn_stats {    /* counters since create (or the last reset) */
    uint64_t kernel_launches;
    uint64_t planes_processed;
    uint64_t h2d_bytes, d2h_bytes;   /* what crossed PCIe: kept rows up, interpolated rows down */
    uint64_t frames;
    uint64_t host_copy_bytes;        /* rows copied by host threads: kept field + border row source -> destination (the
                                        reference's BitBlt, :361-391), copied planes, staging of pageable buffers */
} sn_stats;

/* Create a context on cfg->device. Returns SN_OK and *out, or an error (message via
 * sangnom_cuda_last_error(NULL)). */
SN_API int sangnom_cuda_create(const sn_config* cfg, sn_ctx** out);
SN_API void sangnom_cuda_destroy(sn_ctx* ctx);

/* HOST buffers. Pinned (cudaHostAlloc / cudaHostRegister / sangnom_cuda_host_pin) buffers are DMA'd directly;
 * pageable buffers go through the context's pinned staging. Only the kept field goes up and only the interpolated
 * rows come down; the kept rows and the border row of dst are copied src -> dst on the host meanwhile (nothing at all
 * when dst already holds them: src == dst, SN_MODE_FIELD). Jobs may be in any order; they are grouped by `frame`
 * and run in plane order. Synchronous: returns when every dst is complete. Frames are pipelined internally (host
 * copies | H2D | kernels | D2H, four chunks of frames in flight per device, one host thread per device). */
SN_API int sangnom_cuda_process_planes(sn_ctx* ctx, const sn_plane_job* jobs, int njobs);

/* The same work split in two calls so that consecutive batches overlap (batch k+1 uploads while batch k
 * downloads): submit validates and queues the batch and returns a ticket without waiting for any device work; wait
 * returns when every dst of that batch (and of all earlier ones) is complete, or the first error any of them met
 * (a failed batch does not disturb the others). The job array is copied; the src/dst BUFFERS must stay valid and
 * untouched until wait returns. process_planes(jobs) == submit(jobs) + wait(ticket). */
typedef uint64_t sn_ticket;
SN_API int sangnom_cuda_submit(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, sn_ticket* ticket);
SN_API int sangnom_cuda_wait(sn_ctx* ctx, sn_ticket ticket);

/* DEVICE buffers: src/dst are device pointers on ctx's device. Asynchronous on `cuda_stream`, a
 * cudaStream_t passed as void* (NULL is CUDA's legacy default stream, as everywhere in CUDA;
 * SN_STREAM_CONTEXT selects the context's own compute stream); no host/device copies.
 * The job list is consumed before the call returns. Calls on one context share its cost-state
 * scratch, so they must be stream-ordered with respect to each other. */
#define SN_STREAM_CONTEXT ((void*)(intptr_t)-1)
SN_API int sangnom_cuda_process_planes_device(sn_ctx* ctx, const sn_plane_job* jobs, int njobs, void* cuda_stream);

/* Block until everything queued by _device on the context's own stream has finished. */
SN_API int sangnom_cuda_synchronize(sn_ctx* ctx);

/* aa/aac (0..128) -> threshold in sample units, float arithmetic as SangNom2.cpp:280-282. */
SN_API float sangnom_cuda_threshold(int aa, int bits_per_component, int sample_type);

SN_API int sangnom_cuda_get_limits(int device, sn_limits* out);
SN_API int sangnom_cuda_get_stats(sn_ctx* ctx, sn_stats* out);
SN_API void sangnom_cuda_reset_stats(sn_ctx* ctx);

/* Pinned host memory helpers for callers that want the zero-staging path. */
SN_API void* sangnom_cuda_host_alloc(size_t bytes);
SN_API void sangnom_cuda_host_free(void* p);
/* Pin a long-lived host range the caller owns (a frame server's recycled frame buffers) so that planes inside it are
 * DMA'd directly (cudaHostRegister). The range must stay allocated until it is unpinned; everything still pinned is
 * released by sangnom_cuda_destroy. Pinning costs about a millisecond per few MB: worth it only for buffers that
 * come back. Returns SN_OK, or SN_ERR_CUDA when the driver refuses (the range then simply stays pageable). */
SN_API int sangnom_cuda_host_pin(sn_ctx* ctx, void* base, size_t bytes);
SN_API int sangnom_cuda_host_unpin(sn_ctx* ctx, void* base);
/* Number of devices (pipelines) behind the context. */
SN_API int sangnom_cuda_device_count(sn_ctx* ctx);

/* ---- Anti-aliasing chain: SangNom2(dh=true) -> turn -> SangNom2(dh=true) -> turn back, on the device ----------
 * What scripts build from four filters around the reference (README.md:43-46 `dh`; AviSynth TurnRight/TurnLeft or
 * VapourSynth std.Transpose between the two calls): a W x H frame becomes 2W x 2H. Here the frame is uploaded once,
 * stays in device memory across both passes and both turns, and is downloaded once. Each pass is bit-identical to
 * a stand-alone SangNom2(dh=true) call on the (turned) clip; planes of a frame share a pool per pass as usual. */
typedef struct sn_chain sn_chain;

#define SN_TURN_TRANSPOSE  0   /* transpose, pass, transpose */
#define SN_TURN_RIGHT_LEFT 1   /* TurnRight (clockwise), pass, TurnLeft */
#define SN_TURN_LEFT_RIGHT 2   /* TurnLeft, pass, TurnRight */

typedef struct sn_chain_config {
    int abi_version;          /* SANGNOM_CUDA_ABI_VERSION */
    int device;
    int sample_type;          /* SN_SAMPLE_* */
    int width, height;        /* INPUT luma size */
    int turn;                 /* SN_TURN_* */
    int max_frames_in_flight; /* 0 = default */
    int flags;                /* reserved, 0 */
} sn_chain_config;

typedef struct sn_chain_job {
    const void* src;          /* host, W x H samples of one plane */
    ptrdiff_t src_pitch;      /* bytes */
    void* dst;                /* host, 2W x 2H samples */
    ptrdiff_t dst_pitch;
    int width, height;        /* W, H of this plane */
    int offset1, offset2;     /* field offset (0 keep as top rows, 1 as bottom rows) of pass 1 / pass 2 - what
                                 order/parity resolve to in the two SangNom2 calls */
    float threshold;          /* sangnom_cuda_threshold(aa or aac, ...) - both passes */
    int plane;                /* 0 Y, 1 U, 2 V */
    int frame;                /* caller's frame key */
} sn_chain_job;

typedef struct sn_chain_stats {
    uint64_t kernel_launches;       /* turn kernels */
    uint64_t pass_kernel_launches;  /* row-sweep kernels of both passes */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t frames;
} sn_chain_stats;

SN_API int sangnom_cuda_chain_create(const sn_chain_config* cfg, sn_chain** out);
SN_API void sangnom_cuda_chain_destroy(sn_chain* chain);
/* Synchronous; frames are pipelined internally (upload | passes + turns | download). */
SN_API int sangnom_cuda_chain_process(sn_chain* chain, const sn_chain_job* jobs, int njobs);
SN_API int sangnom_cuda_chain_get_stats(sn_chain* chain, sn_chain_stats* out);
SN_API const char* sangnom_cuda_chain_last_error(sn_chain* chain);

/* The turn on its own, device planes: dst (height x width samples) = src (width x height) transposed (kind 0),
 * turned clockwise (1) or counter-clockwise (2). Asynchronous on `cuda_stream` (a cudaStream_t as void*). */
/* sn_turn_plane.flags: the caller does not care about the bytes between the end of a dst row (`height` samples) and
 * the next multiple of 16 bytes inside its pitch - they may be overwritten. Lets planes with any row length take the
 * tensor-map (TMA) path, whose stores clip at 16-byte granularity; without it only planes whose dst rows are a whole
 * number of 16-byte pieces do, the others take the plain kernel. */
#define SN_TURN_DST_PADDING_WRITABLE 1
typedef struct sn_turn_plane {
    const void* src; ptrdiff_t src_pitch;
    void* dst; ptrdiff_t dst_pitch;
    int width, height;        /* of src, in samples */
    int flags;                /* SN_TURN_* bits */
} sn_turn_plane;
SN_API int sangnom_cuda_turn_planes_device(int sample_type, int kind, const sn_turn_plane* planes, int nplanes, void* cuda_stream);

/* Last error text of ctx (or of the calling thread's last failed create when ctx == NULL). */
SN_API const char* sangnom_cuda_last_error(sn_ctx* ctx);
SN_API int sangnom_cuda_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SANGNOM_CUDA_H */
