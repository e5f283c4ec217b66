// Shared declarations of the TMA-streamed decode kernels (gemv_stream.cu,
// attention stage, engine): device weight layouts, prologue descriptor, PTX
// helpers for mbarrier / cp.async.bulk / programmatic dependent launch.
#pragma once
#include "zb_common.cuh"
#include "zb_quant.cuh"

namespace zb {

// ---- PTX helpers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a byte-count mismatch would otherwise hang the GPU; trap instead (surfaces as a launch failure).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
        if (spins > 4) __nanosleep(32);  // long waits (a whole pipeline stage) must not steal issue slots from the working warps
    }
}
// 1-D bulk copy global -> shared through the TMA engine (UBLKCP): dst/src 16-B aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
// Programmatic dependent launch: let the next kernel in the stream start its
// weight prefetch now / wait until the previous kernel's results are visible.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// packed f32x2 arithmetic (FADD2 / FFMA2 on sm_100a): halves the issue slots of the dot products
__device__ __forceinline__ uint64_t pack2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t a, uint32_t b) {
    uint64_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ float sum2(uint64_t v) {
    float a, b;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a + b;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- stream layouts ----------------------------------------------------------
// Every matrix the streamed kernels read is stored as
//   main: rows x row_main bytes, each row 16-B aligned (TMA bulk-copy friendly)
//   aux : rows x row_aux bytes of fp16 block scales that would break that alignment
// Q4_0: main = 16-B nibble blocks,  aux = fp16 d per 32   (the reference's own GPU layout, gemm_q4.h:3-4)
// Q8_0: main = 32-B int8 blocks,    aux = fp16 d per 32
// Q4_K: main = raw 144-B super-blocks (already 16-B multiples), no aux
// Q5_K: main = raw 176-B super-blocks, no aux
// Q6_K: main = ql[128] qh[64] sc[16] = 208 B, aux = fp16 d per 256
// All are pure byte moves of the GGUF blocks: dequantised values stay bit-exact.
__host__ __device__ inline int stream_main_per8(int t) {  // main bytes per 8 chunks (256 weights)
    switch (t) {
        case kQ4_0: return 128;
        case kQ8_0: return 256;
        case kQ4_K: return 144;
        case kQ5_K: return 176;
        case kQ6_K: return 208;
    }
    return 0;
}
__host__ __device__ inline bool stream_is_kquant(int t) { return t == kQ4_K || t == kQ5_K || t == kQ6_K; }
// bytes of main for `chunks` 32-weight chunks starting at a chunk index that is a multiple of 8 for K-quants
__host__ __device__ inline int stream_main_bytes(int t, int chunks) {
    return stream_is_kquant(t) ? (chunks >> 3) * stream_main_per8(t) : chunks * (stream_main_per8(t) >> 3);
}
__host__ __device__ inline int stream_aux_bytes(int t, int chunks) {
    if (t == kQ4_0 || t == kQ8_0) return chunks * 2;
    if (t == kQ6_K) return (chunks >> 3) * 2;
    return 0;
}

struct StreamW {
    const uint8_t* main;
    const uint8_t* aux;
    int type, M, K;
};

// What a streamed GEMV does to produce its activation vector x[K] before the
// contraction (every CTA recomputes it; it is K floats):
//   v = a                                  (or sum_k mix_w[k]*a[k*mix_stride + i], the MoE combine, moe.go:470-479)
//   if w1:  v = rmsnorm(v, w1)            (Gemma-3 post-attention / post-FFN norm)
//   if r:   v = v + r ; CTA 0 stores v to sum_out   (the residual stream)
//   if w2:  x = rmsnorm(v, w2) else x = v
//   swiglu: x[i] = silu(a[i]) * a[K + i]   (exclusive with the above)
struct Prologue {
    const float* a;
    const float* r;
    const float* w1;
    const float* w2;
    float* sum_out;
    const float* mix_w;
    int mix_n, mix_stride;
    float eps;
    int swiglu;
    int a_rep, a_rep_stride;   // CTA c reads a + (c % a_rep) * a_rep_stride: identical copies spread the L2 hot spot of a vector every SM reads
};

}  // namespace zb
