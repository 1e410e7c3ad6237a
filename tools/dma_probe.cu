// Probe: how fast do the copy engines move the kept / interpolated rows of a batch of frames (every other row of every
// plane, frames back to back in one pinned arena) when issued (A) as one 2-D copy per plane, (B) as one 3-D copy per
// plane kind per chunk of frames (depth = frames). Both directions at once, as the host pipeline runs them.
//   nvcc -O2 -o tools/dma_probe tools/dma_probe.cu && tools/dma_probe
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Plane { size_t off; int row, rows; };   // offset inside a frame, row bytes, rows

int main()
{
    const int W = 1920, H = 1080, frames = 592, chunk = 37;
    const Plane planes[3] = { { 0, W, H }, { (size_t)W * H, W / 2, H / 2 }, { (size_t)W * H + (size_t)W * H / 4, W / 2, H / 2 } };
    const size_t frame_bytes = (size_t)W * H * 3 / 2;
    char *hsrc, *hdst, *dsrc, *dout;
    CK(cudaHostAlloc(&hsrc, frames * frame_bytes, cudaHostAllocPortable));
    CK(cudaHostAlloc(&hdst, frames * frame_bytes, cudaHostAllocPortable));
    CK(cudaMalloc(&dsrc, frames * frame_bytes / 2 + (1 << 20)));
    CK(cudaMalloc(&dout, frames * frame_bytes / 2 + (1 << 20)));
    cudaStream_t up[4], down[4];
    for (int i = 0; i < 4; ++i) { CK(cudaStreamCreateWithFlags(&up[i], cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&down[i], cudaStreamNonBlocking)); }
    const size_t half = frame_bytes / 2;
    auto run = [&](int mode, bool do_up, bool do_down, int nstreams) {
        auto t0 = std::chrono::steady_clock::now();
        for (int rep = 0; rep < 3; ++rep)
            for (int c = 0; c * chunk < frames; ++c) {
                cudaStream_t su = up[c % nstreams], sd = down[c % nstreams];
                const int f0 = c * chunk;
                if (mode == 0) {
                    for (int f = f0; f < f0 + chunk; ++f)
                        for (int p = 0; p < 3; ++p) {
                            const Plane& pl = planes[p];
                            // device layout: per frame, the planes' kept rows back to back
                            char* d = dsrc + (size_t)f * half + pl.off / 2;
                            if (do_up) CK(cudaMemcpy2DAsync(d, pl.row, hsrc + f * frame_bytes + pl.off, 2 * pl.row, pl.row, pl.rows / 2, cudaMemcpyHostToDevice, su));
                            char* o = dout + (size_t)f * half + pl.off / 2;
                            if (do_down) CK(cudaMemcpy2DAsync(hdst + f * frame_bytes + pl.off + pl.row, 2 * pl.row, o, pl.row, pl.row, pl.rows / 2 - 1, cudaMemcpyDeviceToHost, sd));
                        }
                } else {
                    // device layout: plane-major inside the chunk, so that a slice is exactly rows/2 rows
                    size_t doff = (size_t)f0 * half;
                    for (int p = 0; p < 3; ++p) {
                        const Plane& pl = planes[p];
                        cudaMemcpy3DParms q = {};
                        q.srcPtr = make_cudaPitchedPtr(hsrc + f0 * frame_bytes + pl.off, 2 * pl.row, pl.row, frame_bytes / (2 * pl.row));
                        q.dstPtr = make_cudaPitchedPtr(dsrc + doff, pl.row, pl.row, pl.rows / 2);
                        q.extent = make_cudaExtent(pl.row, pl.rows / 2, chunk);
                        q.kind = cudaMemcpyHostToDevice;
                        if (do_up) CK(cudaMemcpy3DAsync(&q, su));
                        cudaMemcpy3DParms r = {};
                        r.srcPtr = make_cudaPitchedPtr(dout + doff, pl.row, pl.row, pl.rows / 2);
                        r.dstPtr = make_cudaPitchedPtr(hdst + f0 * frame_bytes + pl.off + pl.row, 2 * pl.row, pl.row, frame_bytes / (2 * pl.row));
                        r.extent = make_cudaExtent(pl.row, pl.rows / 2 - 1, chunk);
                        r.kind = cudaMemcpyDeviceToHost;
                        if (do_down) CK(cudaMemcpy3DAsync(&r, sd));
                        doff += (size_t)pl.row * (pl.rows / 2) * chunk;
                    }
                }
            }
        CK(cudaDeviceSynchronize());
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return 3.0 * frames * half / dt / 1e9;
    };
    for (int nstreams : { 1, 4 })
        for (int mode = 0; mode < 2; ++mode) {
            run(mode, true, true, nstreams);
            printf("%s, %d stream(s) per direction: up alone %.1f  down alone %.1f  both %.1f GB/s per direction\n", mode ? "3-D copy per plane kind per chunk" : "2-D copy per plane",
                   nstreams, run(mode, true, false, nstreams), run(mode, false, true, nstreams), run(mode, true, true, nstreams));
        }
    // contiguous reference
    {
        auto t0 = std::chrono::steady_clock::now();
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemcpyAsync(dsrc, hsrc, frames * half, cudaMemcpyHostToDevice, up[0]));
            CK(cudaMemcpyAsync(hdst, dout, frames * half, cudaMemcpyDeviceToHost, down[0]));
        }
        CK(cudaDeviceSynchronize());
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("contiguous, both directions: %.1f GB/s per direction\n", 3.0 * frames * half / dt / 1e9);
    }
    return 0;
}
