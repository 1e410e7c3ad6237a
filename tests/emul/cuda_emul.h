// Host emulation of the handful of CUDA constructs the SangNom kernels use, so that the SAME kernel
// source (csrc/*.cuh) can be compiled with g++ and run block-by-block on CPU threads.
// TEST INFRASTRUCTURE ONLY (tests/test_kernel_emulation.py): it lets the `-m "not gpu"` suite check
// the packed-lane arithmetic, the halo exchange and the cost-state hand-over of the device code
// against the oracle without a GPU. It is not a CPU fallback: nothing in the product links it.
#pragma once
#include <algorithm>
#include <barrier>
#include <math.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define SN_HOST_EMULATION 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__
#define __align__(x) alignas(x)

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct dim3e { unsigned x = 0, y = 0, z = 0; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{ x, y }; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{ x, y, z, w }; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }

typedef int cudaError_t;
typedef void* cudaStream_t;

namespace emul {
inline thread_local dim3e tidx, bidx, bdim;
inline thread_local unsigned char* smem = nullptr;
inline thread_local std::barrier<>* bar = nullptr;
// thread block cluster: rank of this block, cluster size, every block's shared memory, cluster-wide barrier
inline thread_local unsigned crank = 0, csize = 1;
inline thread_local unsigned char* const* cluster_smem = nullptr;
inline thread_local std::barrier<>* cluster_bar = nullptr;
}  // namespace emul
#define threadIdx (emul::tidx)
#define blockIdx (emul::bidx)
#define blockDim (emul::bdim)
#define SN_DYNAMIC_SMEM(name) unsigned char* name = emul::smem

inline void __syncthreads() { emul::bar->arrive_and_wait(); }
template <typename T> inline T __ldg(const T* p) { return *p; }

inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift)
{
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (shift & 31));
}
inline uint32_t emul_prmt(uint32_t a, uint32_t b, uint32_t s)   // PTX prmt.b32, default mode
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) { return emul_prmt(a, b, s & 0x7777); }
inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) { const int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF; r |= (uint32_t)std::abs(x - y) << (8 * i); }
    return r;
}
inline uint32_t __vmaxu2(uint32_t a, uint32_t b) { return std::max(a & 0xFFFFu, b & 0xFFFFu) | (std::max(a >> 16, b >> 16) << 16); }
inline uint32_t __vminu2(uint32_t a, uint32_t b) { return std::min(a & 0xFFFFu, b & 0xFFFFu) | (std::min(a >> 16, b >> 16) << 16); }
inline uint32_t __vimin3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vminu2(__vminu2(a, b), c); }
inline uint32_t __vimin3_u32(uint32_t a, uint32_t b, uint32_t c) { return std::min(std::min(a, b), c); }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
using std::max;


using std::abs;
using std::min;

namespace emul {
// Run one block of `threads` emulated threads of `kernel(args...)` with `smem_bytes` of shared memory.
template <typename F>
void run_block(unsigned block, unsigned threads, size_t smem_bytes, F&& body)
{
    std::vector<unsigned char> shared(smem_bytes + 64, 0xA5);     // poisoned: reads of unwritten smem show up
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(shared.data()) + 15) & ~(uintptr_t)15);
    std::barrier<> barrier((std::ptrdiff_t)threads);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            tidx = dim3e{ t, 0, 0 }; bidx = dim3e{ block, 0, 0 }; bdim = dim3e{ threads, 1, 1 };
            smem = sm; bar = &barrier;
            body();
        });
    for (auto& th : pool) th.join();
}

// Run one cluster of `nblocks` blocks (block indices first_block..first_block+nblocks-1) concurrently.
template <typename F>
void run_cluster(unsigned first_block, unsigned nblocks, unsigned threads, size_t smem_bytes, F&& body)
{
    std::vector<std::vector<unsigned char>> shared(nblocks, std::vector<unsigned char>(smem_bytes + 64, 0xA5));
    std::vector<unsigned char*> bases(nblocks);
    for (unsigned b = 0; b < nblocks; ++b)
        bases[b] = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(shared[b].data()) + 15) & ~(uintptr_t)15);
    std::vector<std::unique_ptr<std::barrier<>>> block_bars;
    for (unsigned b = 0; b < nblocks; ++b) block_bars.emplace_back(new std::barrier<>((std::ptrdiff_t)threads));
    std::barrier<> cbar((std::ptrdiff_t)threads * nblocks);
    std::vector<std::thread> pool;
    for (unsigned b = 0; b < nblocks; ++b)
        for (unsigned t = 0; t < threads; ++t)
            pool.emplace_back([&, b, t] {
                tidx = dim3e{ t, 0, 0 }; bidx = dim3e{ first_block + b, 0, 0 }; bdim = dim3e{ threads, 1, 1 };
                smem = bases[b]; bar = block_bars[b].get();
                crank = b; csize = nblocks; cluster_smem = bases.data(); cluster_bar = &cbar;
                body();
            });
    for (auto& th : pool) th.join();
}
}  // namespace emul
