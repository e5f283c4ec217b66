// Fused activation prologue of the streamed decode GEMVs: builds x[K] in shared memory (RMSNorm / Add+RMSNorm /
// Norm+Add+Norm / SwiGLU / MoE combine) in the unit-major padded order the dot products consume it.  Shared by the
// CUDA-core kernel (gemv_stream.cu) and the tensor-core kernel (gemv_mma.cu).
#pragma once
#include "zb_stream.cuh"

namespace zb {
// A lane works on one "unit" of a row at a time: a 32-weight block (Q4_0, Q8_0) or a 64-weight group of a
// K-quant super-block (both sub-blocks that share 32 bytes of nibbles / one pair of scales).
__host__ __device__ constexpr int unit_w(int t) { return (t == kQ4_K || t == kQ5_K || t == kQ6_K) ? 64 : 32; }

// position of element k of x inside shared memory: unit-major in the order the format's dot product
// consumes it; units are padded by 16 bytes so the lanes' 128-bit reads are bank-conflict free without address math.
template <int TYPE>
__device__ __forceinline__ int xpos(int k) {
    int u, p;
    if (TYPE == kQ4_K || TYPE == kQ5_K) {        // unit = group g of a super-block: [32 low-nibble | 32 high-nibble] weights
        int b = k >> 8, e = k & 255;
        u = (b << 2) + (e >> 6);
        p = (((e >> 5) & 1) << 5) + (e & 31);
    } else if (TYPE == kQ6_K) {                  // unit = (half, 16-lane half of l): q1 | q2 | q3 | q4, 16 each
        int b = k >> 8, e = k & 255, l = e & 31;
        u = (b << 2) + ((e >> 7) << 1) + (l >> 4);
        p = (((e >> 5) & 3) << 4) + (l & 15);
    } else {
        u = k >> 5;
        p = k & 31;
    }
    return u * (unit_w(TYPE) + 4) + p;  // unit stride padded by one 16-B group: consecutive lanes hit distinct bank groups
}

__device__ __forceinline__ float inv_rms(float sumsq, int D, float eps) {
    return (float)(1.0 / sqrt((double)(sumsq / (float)D + eps)));  // rmsnorm_generic.go:17 (f64 sqrt, one rounding)
}

// The activation vector is built with 128-bit accesses, 8 independent loads in flight per thread per batch:
// a strided scalar loop here costs one L2 round trip per iteration and used to dominate the small GEMVs.
constexpr int kPB = 4;  // float4 loads in flight per thread (per array)

template <int TYPE>
__device__ __forceinline__ float4* xslot(float* xs, int i4) { return reinterpret_cast<float4*>(xs + xpos<TYPE>(i4 << 2)); }

__device__ __forceinline__ float sq4(float4 v, float ss) {
    ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss);
    return fmaf(v.w, v.w, ss);
}

// Four consecutive elements of an exchange slot ((value, epoch) pairs written by a peer over NVLink): spin until all four
// carry this exchange's epoch.  Bounded: a lost peer traps instead of hanging the GPU.
__device__ __forceinline__ float4 ll_load4(const uint2* slot, int i4, unsigned int want) {
    const uint4* p = reinterpret_cast<const uint4*>(slot + 4 * i4);
    uint4 lo, hi;
    unsigned int spins = 0;
    do {
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "l"(p) : "memory");
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(p + 1) : "memory");
        if (++spins > (1u << 24)) __trap();
    } while (lo.y != want || lo.w != want || hi.y != want || hi.w != want);
    return make_float4(__uint_as_float(lo.x), __uint_as_float(lo.z), __uint_as_float(hi.x), __uint_as_float(hi.z));
}

template <int TYPE, int kSThreads>
__device__ void build_x(const Prologue& p, const float* __restrict__ a, int K, float* xs, float4* xsum, float* red, bool lead,
                        unsigned int ll_epoch = 0u) {
    const int tid = threadIdx.x, K4 = K >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    if (p.swiglu) {  // silu_generic.go:22-31: sigmoid in f64, float32(g*sig)*u
        const float4* u4 = reinterpret_cast<const float4*>(a + K);
        for (int base = 0; base < K4; base += kSThreads * kPB) {
            float4 g[kPB], u[kPB];
#pragma unroll
            for (int j = 0; j < kPB; j++) {
                int i = base + tid + kSThreads * j;
                if (i < K4) { g[j] = __ldcg(a4 + i); u[j] = __ldcg(u4 + i); }
            }
#pragma unroll
            for (int j = 0; j < kPB; j++) {
                int i = base + tid + kSThreads * j;
                if (i < K4) {
                    float gg[4] = {g[j].x, g[j].y, g[j].z, g[j].w}, uu[4] = {u[j].x, u[j].y, u[j].z, u[j].w}, o[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        double gv = (double)gg[e];
                        double sig = 1.0 / (1.0 + exp(-gv));
                        o[e] = (float)(gv * sig) * uu[e];
                    }
                    *xslot<TYPE>(xs, i) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    } else {
        const float4* r4 = reinterpret_cast<const float4*>(p.r);
        float4* so4 = (lead && p.sum_out) ? reinterpret_cast<float4*>(p.sum_out) : nullptr;
        const bool add_now = !p.w1 && p.r;
        float ss = 0.0f;
        for (int base = 0; base < K4; base += kSThreads * kPB) {
            float4 v[kPB], r[kPB];
#pragma unroll
            for (int j = 0; j < kPB; j++) {
                int i = base + tid + kSThreads * j;
                if (i < K4) {
                    if (p.mix_n > 0) {  // MoE combine: out = 0; out += y_k * w_k in selection order (moe.go:470-479)
                        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int k = 0; k < p.mix_n; k++) {
                            float4 yk = ll_epoch ? ll_load4(reinterpret_cast<const uint2*>(a) + (size_t)k * p.mix_stride, i, ll_epoch)
                                                 : __ldcg(reinterpret_cast<const float4*>(a + (size_t)k * p.mix_stride) + i);
                            float wk = p.mix_w[k];
                            t.x = t.x + yk.x * wk; t.y = t.y + yk.y * wk; t.z = t.z + yk.z * wk; t.w = t.w + yk.w * wk;
                        }
                        v[j] = t;
                    } else {
                        v[j] = __ldcg(a4 + i);
                    }
                    if (add_now) r[j] = __ldcg(r4 + i);
                }
            }
#pragma unroll
            for (int j = 0; j < kPB; j++) {
                int i = base + tid + kSThreads * j;
                if (i < K4) {
                    float4 t = v[j];
                    if (add_now) {
                        t.x = t.x + r[j].x; t.y = t.y + r[j].y; t.z = t.z + r[j].z; t.w = t.w + r[j].w;
                        if (so4) so4[i] = t;
                    }
                    *xslot<TYPE>(xs, i) = t;
                    ss = sq4(t, ss);
                }
            }
        }
        if (p.w1) {
            const float4* w4 = reinterpret_cast<const float4*>(p.w1);
            float s1 = inv_rms(block_sum(ss, red), K, p.eps);
            ss = 0.0f;
            for (int base = 0; base < K4; base += kSThreads * kPB) {
                float4 w[kPB], r[kPB];
#pragma unroll
                for (int j = 0; j < kPB; j++) {
                    int i = base + tid + kSThreads * j;
                    if (i < K4) {
                        w[j] = __ldg(w4 + i);
                        if (p.r) r[j] = __ldcg(r4 + i);
                    }
                }
#pragma unroll
                for (int j = 0; j < kPB; j++) {
                    int i = base + tid + kSThreads * j;
                    if (i < K4) {
                        float4* at = xslot<TYPE>(xs, i);
                        float4 t = *at;
                        t.x = t.x * s1 * w[j].x; t.y = t.y * s1 * w[j].y; t.z = t.z * s1 * w[j].z; t.w = t.w * s1 * w[j].w;
                        if (p.r) {
                            t.x = t.x + r[j].x; t.y = t.y + r[j].y; t.z = t.z + r[j].z; t.w = t.w + r[j].w;
                            if (so4) so4[i] = t;
                        }
                        *at = t;
                        ss = sq4(t, ss);
                    }
                }
            }
        }
        if (p.w2) {
            const float4* w4 = reinterpret_cast<const float4*>(p.w2);
            float s2 = inv_rms(block_sum(ss, red), K, p.eps);
            for (int base = 0; base < K4; base += kSThreads * kPB) {
                float4 w[kPB];
#pragma unroll
                for (int j = 0; j < kPB; j++) {
                    int i = base + tid + kSThreads * j;
                    if (i < K4) w[j] = __ldg(w4 + i);
                }
#pragma unroll
                for (int j = 0; j < kPB; j++) {
                    int i = base + tid + kSThreads * j;
                    if (i < K4) {
                        float4* at = xslot<TYPE>(xs, i);
                        float4 t = *at;
                        t.x = t.x * s2 * w[j].x; t.y = t.y * s2 * w[j].y; t.z = t.z * s2 * w[j].z; t.w = t.w * s2 * w[j].w;
                        *at = t;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (TYPE == kQ4_K || TYPE == kQ5_K) {
        // per-unit partial sums of x, one per scale group of the unit (they carry the dmin term and the float-trick bias):
        // Q4_K / Q5_K: (low-nibble 32, high-nibble 32); Q6_K: (q1, q2, q3, q4) 16 each; Q4_0: (all 32)
        constexpr int UWP = unit_w(TYPE) + 4, NG = unit_w(TYPE) / 4;  // 16-B groups per unit
        const int U = K / unit_w(TYPE);
        for (int u = tid; u < U; u += kSThreads) {
            const float4* xp = reinterpret_cast<const float4*>(xs + u * UWP);
            float sum[4] = {0.f, 0.f, 0.f, 0.f};
            constexpr int per = TYPE == kQ6_K ? 4 : (TYPE == kQ4_0 ? 8 : 8);  // groups per partial sum
#pragma unroll
            for (int j = 0; j < NG; j++) {
                float4 t = xp[j];
                sum[j / per] += (t.x + t.y) + (t.z + t.w);
            }
            xsum[u] = make_float4(sum[0], sum[1], sum[2], sum[3]);
        }
        __syncthreads();
    }
}

}  // namespace zb
