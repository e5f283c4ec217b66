// Shared device helpers for the sm_100a kernel library.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

#include <mutex>
#include <thread>
#include <vector>

#define ZB_API extern "C" __attribute__((visibility("default")))

#define ZB_WARP 32
#define ZB_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

namespace zb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `red` holds 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();  // protect `red` against a previous use
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : 0.0f;
    return warp_sum(t);
}

__device__ __forceinline__ float h2f(uint16_t bits) { return __half2float(__ushort_as_half(bits)); }

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// Streaming 128-bit load: weights are read exactly once per token, keep them out of L1.
__device__ __forceinline__ uint4 ldg128_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg32_stream(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint16_t ldg16(const void* p) { return __ldg(reinterpret_cast<const uint16_t*>(p)); }

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Host-side repacks are byte permutations over independent row tiles / blocks: split them over the host cores (a 42 GB model
// otherwise spends a minute in single-threaded memcpy loops).  ZB_HOST_THREADS overrides; under torchrun the cores are
// shared by WORLD_SIZE loaders.
template <class F>
inline void zb_parallel_for(int64_t n, int64_t min_per_thread, F fn) {
    int threads = (int)std::thread::hardware_concurrency();
    if (const char* w = getenv("WORLD_SIZE")) { int ws = atoi(w); if (ws > 1) threads /= ws; }
    if (const char* t = getenv("ZB_HOST_THREADS")) { int v = atoi(t); if (v > 0) threads = v; }
    if (threads > 32) threads = 32;
    if (min_per_thread < 1) min_per_thread = 1;
    if ((int64_t)threads > n / min_per_thread) threads = (int)(n / min_per_thread);
    if (threads <= 1) { for (int64_t i = 0; i < n; i++) fn(i); return; }
    std::vector<std::thread> pool;
    const int64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        const int64_t lo = t * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        pool.emplace_back([lo, hi, &fn] { for (int64_t i = lo; i < hi; i++) fn(i); });
    }
    for (auto& th : pool) th.join();
}

// Function attributes (opt-in dynamic shared memory) belong to the device / context, not to the process: one process may
// drive engines on several GPUs (zb_engine_opts.device), from several threads.  Per-device, mutex-guarded "configured up to
// `need`" state for a launcher; `fn` does the cudaFuncSetAttribute calls.
struct DeviceOnce {
    std::mutex mu;
    size_t val[64] = {};
    template <class F>
    cudaError_t ensure(size_t need, F fn) {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) d = 0;
        d &= 63;
        std::lock_guard<std::mutex> lock(mu);
        if (val[d] >= need) return cudaSuccess;
        cudaError_t e = fn();
        if (e == cudaSuccess) val[d] = need;
        return e;
    }
};

}  // namespace zb
