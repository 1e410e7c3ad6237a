// Probe: bandwidth of SM-issued copies over PCIe (mapped pinned host memory) against the copy engines,
// for the access patterns of the host path: rows of `row` bytes, every other row (the kept field).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/zerocopy_bw tools/zerocopy_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// copy `rows` rows of `row16` uint4 each; src/dst strides in uint4 units
__global__ void copy_rows(const uint4* __restrict__ src, size_t sstride, uint4* __restrict__ dst, size_t dstride, int rows, int row16, int unroll_dummy)
{
    const size_t total = (size_t)rows * row16;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 4 independent 16-byte loads in flight per thread
    for (; i + 3 * step < total; i += 4 * step) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const size_t j = i + k * step; const size_t r = j / row16, c = j % row16; v[k] = src[r * sstride + c]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const size_t j = i + k * step; const size_t r = j / row16, c = j % row16; dst[r * dstride + c] = v[k]; }
    }
    for (; i < total; i += step) { const size_t r = i / row16, c = i % row16; dst[r * dstride + c] = src[r * sstride + c]; }
}

static float time_ms(cudaStream_t s, void (*fn)(void*), void* arg, int iters)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    fn(arg); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a, s));
    for (int i = 0; i < iters; ++i) fn(arg);
    CK(cudaEventRecord(b, s));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

struct Args { const uint4* src; size_t ss; uint4* dst; size_t ds; int rows, row16, blocks, threads; cudaStream_t st; };
static void run(void* p) { Args* a = (Args*)p; copy_rows<<<a->blocks, a->threads, 0, a->st>>>(a->src, a->ss, a->dst, a->ds, a->rows, a->row16, 0); }

int main()
{
    const int row = 1920, rows = 270 * 1024;                  // 0.53 GB of kept rows
    const size_t field = (size_t)row * rows, full = 2 * field;
    uint4 *h_src, *h_dst, *d_a, *d_b;
    CK(cudaHostAlloc((void**)&h_src, full, cudaHostAllocDefault));
    CK(cudaHostAlloc((void**)&h_dst, full, cudaHostAllocDefault));
    memset(h_src, 1, full); memset(h_dst, 2, full);
    CK(cudaMalloc((void**)&d_a, full)); CK(cudaMalloc((void**)&d_b, full));
    CK(cudaMemset(d_a, 3, full)); CK(cudaMemset(d_b, 4, full));
    uint4 *hd_src, *hd_dst;
    CK(cudaHostGetDevicePointer((void**)&hd_src, h_src, 0)); CK(cudaHostGetDevicePointer((void**)&hd_dst, h_dst, 0));
    printf("host ptr == device alias: %d %d\n", (void*)hd_src == (void*)h_src, (void*)hd_dst == (void*)h_dst);
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    const int r16 = row / 16;
    const int blockset[] = { 4, 8, 16, 32, 74, 148, 296 };
    for (int threads : { 256, 512 })
        for (int nb : blockset) {
            Args g{ hd_src, (size_t)2 * r16, d_a, (size_t)r16, rows, r16, nb, threads, s1 };       // gather kept rows (strided host reads)
            const float tg = time_ms(s1, run, &g, 3);
            Args sc{ d_b, (size_t)r16, hd_dst, (size_t)2 * r16, rows, r16, nb, threads, s1 };      // scatter rows (strided host writes)
            const float ts = time_ms(s1, run, &sc, 3);
            Args sf{ d_b, (size_t)r16, hd_dst, (size_t)r16, 2 * rows, r16, nb, threads, s1 };      // contiguous host writes, full frame
            const float tf = time_ms(s1, run, &sf, 3);
            printf("threads %3d blocks %3d: gather(strided H->D) %6.1f GB/s   scatter(strided D->H) %6.1f GB/s   write full D->H %6.1f GB/s\n",
                   threads, nb, field / tg / 1e6, field / ts / 1e6, full / tf / 1e6);
        }
    // both directions at once, 32 blocks each
    for (int nb : { 16, 32, 74 }) {
        Args g{ hd_src, (size_t)2 * r16, d_a, (size_t)r16, rows, r16, nb, 512, s1 };
        Args sc{ d_b, (size_t)r16, hd_dst, (size_t)2 * r16, rows, r16, nb, 512, s2 };
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        run(&g); run(&sc); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, 0));
        for (int i = 0; i < 3; ++i) { run(&g); run(&sc); }
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(b, 0)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        printf("concurrent gather+scatter, %d blocks each: %6.1f GB/s per direction\n", nb, field * 3 / ms / 1e6);
    }
    // copy engines for comparison: contiguous, and many 2 MB copies
    {
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        CK(cudaMemcpyAsync(d_a, h_src, full, cudaMemcpyHostToDevice, s1)); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, s1)); CK(cudaMemcpyAsync(d_a, h_src, full, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(b, s1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        printf("copy engine H->D contiguous: %6.1f GB/s\n", full / ms / 1e6);
        for (size_t piece : { (size_t)518400, (size_t)2073600, (size_t)3110400 }) {
            const size_t n = full / piece;
            CK(cudaEventRecord(a, s1));
            for (size_t i = 0; i < n; ++i) CK(cudaMemcpyAsync((char*)d_a + i * piece, (char*)h_src + i * piece, piece, cudaMemcpyHostToDevice, s1));
            CK(cudaEventRecord(b, s1)); CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms, a, b));
            printf("copy engine H->D in %zu-byte pieces: %6.1f GB/s\n", piece, n * piece / ms / 1e6);
            CK(cudaEventRecord(a, s2));
            for (size_t i = 0; i < n; ++i) CK(cudaMemcpyAsync((char*)h_dst + i * piece, (char*)d_b + i * piece, piece, cudaMemcpyDeviceToHost, s2));
            CK(cudaEventRecord(b, s2)); CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms, a, b));
            printf("copy engine D->H in %zu-byte pieces: %6.1f GB/s\n", piece, n * piece / ms / 1e6);
        }
    }
    // copy engines, 2-D (pitched) copies: kept field = every other row
    {
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        float ms;
        for (int rows_per_copy : { 270, 540, 270 * 64, rows }) {
            const int n = rows / rows_per_copy;
            CK(cudaEventRecord(a, s1));
            for (int i = 0; i < n; ++i)
                CK(cudaMemcpy2DAsync((char*)d_a + (size_t)i * rows_per_copy * row, row, (char*)h_src + (size_t)i * rows_per_copy * 2 * row, 2 * row, row, rows_per_copy, cudaMemcpyHostToDevice, s1));
            CK(cudaEventRecord(b, s1)); CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms, a, b));
            printf("copy engine 2-D H->D kept rows, %d rows per copy: %6.1f GB/s\n", rows_per_copy, (size_t)n * rows_per_copy * row / ms / 1e6);
            CK(cudaEventRecord(a, s2));
            for (int i = 0; i < n; ++i)
                CK(cudaMemcpy2DAsync((char*)h_dst + (size_t)i * rows_per_copy * 2 * row, 2 * row, (char*)d_b + (size_t)i * rows_per_copy * row, row, row, rows_per_copy, cudaMemcpyDeviceToHost, s2));
            CK(cudaEventRecord(b, s2)); CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms, a, b));
            printf("copy engine 2-D D->H every other row, %d rows per copy: %6.1f GB/s\n", rows_per_copy, (size_t)n * rows_per_copy * row / ms / 1e6);
        }
        // both directions at once, 2-D, one plane per copy
        const int rpc = 540, n = rows / rpc;
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, 0));
        for (int i = 0; i < n; ++i) {
            CK(cudaMemcpy2DAsync((char*)d_a + (size_t)i * rpc * row, row, (char*)h_src + (size_t)i * rpc * 2 * row, 2 * row, row, rpc, cudaMemcpyHostToDevice, s1));
            CK(cudaMemcpy2DAsync((char*)h_dst + (size_t)i * rpc * 2 * row, 2 * row, (char*)d_b + (size_t)i * rpc * row, row, row, rpc, cudaMemcpyDeviceToHost, s2));
        }
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(b, 0)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("copy engine 2-D both directions at once, 540 rows per copy: %6.1f GB/s per direction\n", (size_t)n * rpc * row / ms / 1e6);
        // SM gather (H->D) with copy-engine contiguous D->H, and copy-engine contiguous H->D with SM scatter
        Args g{ hd_src, (size_t)2 * r16, d_a, (size_t)r16, rows, r16, 16, 512, s1 };
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, 0));
        run(&g); CK(cudaMemcpyAsync(h_dst, d_b, field, cudaMemcpyDeviceToHost, s2));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(b, 0)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("SM gather H->D + copy-engine contiguous D->H at once: %6.1f GB/s per direction\n", field / ms / 1e6);
        Args sc{ d_b, (size_t)r16, hd_dst, (size_t)2 * r16, rows, r16, 16, 512, s2 };
        CK(cudaEventRecord(a, 0));
        run(&sc); CK(cudaMemcpyAsync(d_a, h_src, field, cudaMemcpyHostToDevice, s1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(b, 0)); CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("copy-engine contiguous H->D + SM scatter D->H at once: %6.1f GB/s per direction\n", field / ms / 1e6);
    }
    // host memcpy of the kept field (what a CPU-side BitBlt costs), single thread
    {
        char* tmp = (char*)malloc(full);
        memset(tmp, 0, full);
        timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int r = 0; r < rows; ++r) memcpy((char*)h_dst + (size_t)2 * r * row, (char*)h_src + (size_t)2 * r * row, row);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        printf("host memcpy kept rows pinned->pinned, 1 thread: %6.1f GB/s\n", field / s / 1e9);
        free(tmp);
    }
    return 0;
}
