// Decode-step epilogues and bookkeeping kernels (sm_100a): AddRMSNorm,
// NormAdd, RMSNorm, SwiGLU, QK-norm+RoPE, RoPE, RoPE-select, KV append,
// device counters, argmax, scaled softmax, embedding gather.
//
// All of them are HBM/launch-bound row operations: one CTA per row, 128-bit
// accesses where the row is 16-byte aligned, warp-shuffle reductions, and the
// row held in registers between the reduction and the write-back so it is
// read from HBM exactly once.  Every launcher is asynchronous,
// allocation-free and capture-safe (positions come from device counters).
//
// Replaces (reference file:line):
//   fused_add_rmsnorm_f32   internal/cuda/kernels/fused_add_rmsnorm.cu:70
//   fused_norm_add_f32      internal/cuda/kernels/fused_norm_add.cu:66
//   launch_rmsnorm          internal/cuda/kernels/rmsnorm.cu:75
//   fused_swiglu_f32        internal/cuda/kernels/fused_swiglu.cu:26
//   fused_qk_norm_rope_f32  internal/cuda/kernels/fused_qk_norm_rope.cu:96
//   fused_rope_f32          internal/cuda/kernels/fused_rope.cu:55
//   launch_rope_select      internal/cuda/kernels/rope_select.cu:23
//   launch_offset_memcpy[_fp16]  internal/cuda/kernels/offset_memcpy.cu:23,46
//   launch_increment_counter / launch_reset_counter  counter.cu:21,26
//   launch_argmax           internal/cuda/kernels/argmax.cu:94
//   scaled_softmax_f32      internal/cuda/kernels/scaled_softmax.cu:106
//   launch_gather[_i32]     internal/cuda/kernels/gather.cu:48,57
#include <float.h>

#include "zb_common.cuh"

namespace {

using namespace zb;

constexpr int kRowThreads = 256;
constexpr int kMaxPerThread = 16;  // register-resident rows up to 256*16 = 4096... larger rows re-read

// 1/sqrt(mean + eps) evaluated in f64 and rounded once, as the CPU engine
// does (internal/xblas/rmsnorm_generic.go:17); one thread per row, free.
__device__ __forceinline__ float inv_rms(float sumsq, int D, float eps) {
    return (float)(1.0 / sqrt((double)(sumsq / (float)D + eps)));
}

enum NormMode { kNorm = 0, kAddNorm = 1, kNormAdd = 2 };

// MODE kNorm   : out = x*s*w                       (scales[row] = s if non-null)
// MODE kAddNorm: sum = x + r; out = sum*s*w; sum_out = sum
// MODE kNormAdd: out = x*s*w + r
template <int MODE>
__global__ void __launch_bounds__(kRowThreads) rms_row_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                              const float* __restrict__ w, float* __restrict__ out,
                                                              float* __restrict__ aux, float eps, int D) {
    __shared__ float red[32];
    int64_t row = blockIdx.x;
    const float* xr = x + row * D;
    const float* rr = r ? r + row * D : nullptr;
    float* o = out + row * D;
    float v[kMaxPerThread];
    float ss = 0.0f;
    const bool fits = D <= kRowThreads * kMaxPerThread;
    if (fits) {
#pragma unroll
        for (int j = 0; j < kMaxPerThread; j++) {
            int i = threadIdx.x + j * kRowThreads;
            float t = 0.0f;
            if (i < D) {
                t = xr[i];
                if (MODE == kAddNorm) t += rr[i];
            }
            v[j] = t;
            ss = fmaf(t, t, ss);
        }
    } else {
        for (int i = threadIdx.x; i < D; i += kRowThreads) {
            float t = xr[i];
            if (MODE == kAddNorm) t += rr[i];
            ss = fmaf(t, t, ss);
        }
    }
    ss = block_sum(ss, red);
    float s = inv_rms(ss, D, eps);
    if (MODE == kNorm && aux && threadIdx.x == 0) aux[row] = s;
    if (fits) {
#pragma unroll
        for (int j = 0; j < kMaxPerThread; j++) {
            int i = threadIdx.x + j * kRowThreads;
            if (i < D) {
                float y = v[j] * s * w[i];
                if (MODE == kNormAdd) y += rr[i];
                o[i] = y;
                if (MODE == kAddNorm) aux[row * D + i] = v[j];
            }
        }
    } else {
        for (int i = threadIdx.x; i < D; i += kRowThreads) {
            float t = xr[i];
            if (MODE == kAddNorm) t += rr[i];
            float y = t * s * w[i];
            if (MODE == kNormAdd) y += rr[i];
            o[i] = y;
            if (MODE == kAddNorm) aux[row * D + i] = t;
        }
    }
}

// silu(g)*u evaluated exactly as the CPU engine does
// (internal/xblas/silu_generic.go:22-31): sigmoid in f64, float32(g*sig)*u.
// n is FFN-sized (<= 28672 per token), so the f64 exp is free next to the GEMVs.
__global__ void swiglu_kernel(const float* __restrict__ g, const float* __restrict__ u, float* __restrict__ o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double gv = (double)g[i];
    double sig = 1.0 / (1.0 + exp(-gv));
    o[i] = (float)(gv * sig) * u[i];
}

// One CTA per head: RMSNorm(head, wQ|wK) then half-split RoPE with tail
// pass-through (fused_qk_norm_rope.cu:16-79).  headDim <= 1024.
__global__ void qk_norm_rope_kernel(const float* __restrict__ in, const float* __restrict__ wq, const float* __restrict__ wk,
                                    const float* __restrict__ cs, const float* __restrict__ sn, float* __restrict__ out,
                                    float eps, int hd, int nq, int half) {
    extern __shared__ float xn[];  // [hd]
    __shared__ float red[32];
    int head = blockIdx.x;
    const float* x = in + (int64_t)head * hd;
    float* o = out + (int64_t)head * hd;
    const float* w = head < nq ? wq : wk;
    float ss = 0.0f;
    for (int d = threadIdx.x; d < hd; d += blockDim.x) {
        float t = x[d];
        ss = fmaf(t, t, ss);
    }
    ss = block_sum(ss, red);
    float s = inv_rms(ss, hd, eps);
    for (int d = threadIdx.x; d < hd; d += blockDim.x) xn[d] = x[d] * s * w[d];
    __syncthreads();
    for (int d = threadIdx.x; d < hd; d += blockDim.x) {
        if (d < half) {
            float c = cs[d], sv = sn[d], a = xn[d], b = xn[d + half];
            o[d] = a * c - b * sv;
            o[d + half] = a * sv + b * c;
        } else if (d >= 2 * half) {
            o[d] = xn[d];
        }
    }
}

// fused_rope.cu:13-51: input [batch, seq, hd]; angle row = seq index.
__global__ void rope_kernel(const float* __restrict__ in, const float* __restrict__ cs, const float* __restrict__ sn,
                            float* __restrict__ out, int total, int seq, int hd, int half, int stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int d = i % hd;
    int s = (i / hd) % seq;
    int base = i - d;
    if (d < half) {
        float c = cs[s * stride + d], sv = sn[s * stride + d];
        out[i] = in[i] * c - in[base + d + half] * sv;
    } else if (d < 2 * half) {
        int dd = d - half;
        float c = cs[s * stride + dd], sv = sn[s * stride + dd];
        out[i] = in[base + dd] * sv + in[i] * c;
    } else {
        out[i] = in[i];
    }
}

__global__ void rope_select_kernel(const float* __restrict__ ct, const float* __restrict__ st, float* __restrict__ co,
                                   float* __restrict__ so, const int* __restrict__ counter, int half) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    int64_t off = (int64_t)counter[0] * half + i;
    co[i] = ct[off];
    so[i] = st[off];
}

template <typename T>
__global__ void offset_memcpy_kernel(T* __restrict__ dst, const float* __restrict__ src, const int* __restrict__ counter,
                                     int dim, int max_seq) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    int pos = counter[0];
    if (pos < 0 || pos >= max_seq) return;
    if constexpr (sizeof(T) == 2) dst[(int64_t)pos * dim + i] = __float2half(src[i]);
    else dst[(int64_t)pos * dim + i] = src[i];
}

__global__ void counter_add_kernel(int* c, int delta) { atomicAdd(c, delta); }
__global__ void counter_set_kernel(int* c, int v) { *c = v; }

// ---- argmax: strict '>' with lowest-index tie-break (argmax.cu:40-48) -----
__device__ __forceinline__ void amax_merge(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__device__ __forceinline__ void amax_block(float& v, int& i, float* sv, int* si) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        amax_merge(v, i, ov, oi);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0) { sv[w] = v; si[w] = i; }
    __syncthreads();
    if (w == 0) {
        v = l < nw ? sv[l] : -FLT_MAX;
        i = l < nw ? si[l] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, i, o);
            amax_merge(v, i, ov, oi);
        }
    }
}
__global__ void argmax_stage1(const float* __restrict__ in, float* __restrict__ bv, int* __restrict__ bi, int n) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float v = -FLT_MAX;
    int idx = 0x7fffffff;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) amax_merge(v, idx, in[i], i);
    amax_block(v, idx, sv, si);
    if (threadIdx.x == 0) { bv[blockIdx.x] = v; bi[blockIdx.x] = idx; }
}
__global__ void argmax_stage2(const float* __restrict__ bv, const int* __restrict__ bi, int* __restrict__ result, int nb) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float v = -FLT_MAX;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) amax_merge(v, idx, bv[i], bi[i]);
    amax_block(v, idx, sv, si);
    if (threadIdx.x == 0) result[0] = (idx == 0x7fffffff) ? 0 : idx;
}

// softmax(x*scale) along an axis of length A with outer/inner strides
// (scaled_softmax.cu:13-89): one CTA per (outer, inner) stripe.
__global__ void scaled_softmax_kernel(const float* __restrict__ in, float* __restrict__ out, int inner, int A, float scale) {
    __shared__ float red[32];
    int o = blockIdx.x / inner, in_i = blockIdx.x % inner;
    const float* x = in + (int64_t)o * A * inner + in_i;
    float* y = out + (int64_t)o * A * inner + in_i;
    float mx = -FLT_MAX;
    for (int i = threadIdx.x; i < A; i += blockDim.x) mx = fmaxf(mx, x[(int64_t)i * inner] * scale);
    mx = warp_max(mx);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0) red[w] = mx;
    __syncthreads();
    mx = warp_max(l < nw ? red[l] : -FLT_MAX);
    float sum = 0.0f;
    for (int i = threadIdx.x; i < A; i += blockDim.x) {
        float e = expf(x[(int64_t)i * inner] * scale - mx);
        y[(int64_t)i * inner] = e;
        sum += e;
    }
    sum = block_sum(sum, red);
    float inv = 1.0f / sum;
    for (int i = threadIdx.x; i < A; i += blockDim.x) y[(int64_t)i * inner] *= inv;
}

template <typename I>
__global__ void gather_kernel(const float* __restrict__ table, const I* __restrict__ idx, float* __restrict__ out, int D, int V) {
    int row = blockIdx.x;
    int64_t id = (int64_t)idx[row];
    if (id < 0) id = 0;
    if (id >= V) id = V - 1;  // clamp, as gather.cu:20-23
    const float* src = table + id * D;
    float* dst = out + (int64_t)row * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) dst[c] = src[c];
}

inline float bits(unsigned int b) {
    float f;
    memcpy(&f, &b, 4);
    return f;
}

}  // namespace

ZB_API cudaError_t fused_add_rmsnorm_f32(const float* input, const float* residual, const float* weight, float* normed_out,
                                         float* sum_out, unsigned int eps_bits, int rows, int D, cudaStream_t stream) {
    if (rows <= 0 || D <= 0) return cudaSuccess;
    rms_row_kernel<kAddNorm><<<rows, kRowThreads, 0, stream>>>(input, residual, weight, normed_out, sum_out, bits(eps_bits), D);
    return cudaGetLastError();
}

ZB_API cudaError_t fused_norm_add_f32(const float* input, const float* weight, const float* residual, float* output,
                                      unsigned int eps_bits, int rows, int D, cudaStream_t stream) {
    if (rows <= 0 || D <= 0) return cudaSuccess;
    rms_row_kernel<kNormAdd><<<rows, kRowThreads, 0, stream>>>(input, residual, weight, output, nullptr, bits(eps_bits), D);
    return cudaGetLastError();
}

ZB_API cudaError_t launch_rmsnorm(const float* input, const float* weight, float* output, float* scales, unsigned int eps_bits,
                                  int rows, int D, cudaStream_t stream) {
    if (rows <= 0 || D <= 0) return cudaSuccess;
    rms_row_kernel<kNorm><<<rows, kRowThreads, 0, stream>>>(input, nullptr, weight, output, scales, bits(eps_bits), D);
    return cudaGetLastError();
}

ZB_API cudaError_t fused_swiglu_f32(const float* w1, const float* w3, float* output, int n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    swiglu_kernel<<<zb::cdiv(n, 256), 256, 0, stream>>>(w1, w3, output, n);
    return cudaGetLastError();
}

ZB_API cudaError_t fused_qk_norm_rope_f32(const float* input, const float* weightQ, const float* weightK, const float* cosAngles,
                                          const float* sinAngles, float* output, unsigned int eps_bits, int totalHeads,
                                          int headDim, int numQHeads, int halfRotary, cudaStream_t stream) {
    if (totalHeads <= 0 || headDim <= 0) return cudaSuccess;
    if (2 * halfRotary > headDim) return cudaErrorInvalidValue;
    int threads = headDim >= 256 ? 256 : (headDim >= 128 ? 128 : (headDim >= 64 ? 64 : 32));
    qk_norm_rope_kernel<<<totalHeads, threads, headDim * sizeof(float), stream>>>(input, weightQ, weightK, cosAngles, sinAngles,
                                                                                 output, bits(eps_bits), headDim, numQHeads,
                                                                                 halfRotary);
    return cudaGetLastError();
}

ZB_API cudaError_t fused_rope_f32(const float* input, const float* cos_angles, const float* sin_angles, float* output, int batch,
                                  int seq_len, int head_dim, int half_rotary, int cos_stride, cudaStream_t stream) {
    int total = batch * seq_len * head_dim;
    if (total <= 0) return cudaSuccess;
    rope_kernel<<<zb::cdiv(total, 256), 256, 0, stream>>>(input, cos_angles, sin_angles, output, total, seq_len, head_dim,
                                                          half_rotary, cos_stride);
    return cudaGetLastError();
}

ZB_API cudaError_t launch_rope_select(const float* cos_table, const float* sin_table, float* cos_out, float* sin_out,
                                      const int* counter, int halfRotary, cudaStream_t stream) {
    if (halfRotary <= 0) return cudaSuccess;
    rope_select_kernel<<<zb::cdiv(halfRotary, 256), 256, 0, stream>>>(cos_table, sin_table, cos_out, sin_out, counter, halfRotary);
    return cudaGetLastError();
}

ZB_API cudaError_t launch_offset_memcpy(float* dst, const float* src, const int* counter, int dim, int maxSeqLen,
                                        cudaStream_t stream) {
    if (dim <= 0) return cudaSuccess;
    offset_memcpy_kernel<float><<<zb::cdiv(dim, 256), 256, 0, stream>>>(dst, src, counter, dim, maxSeqLen);
    return cudaGetLastError();
}

ZB_API cudaError_t launch_offset_memcpy_fp16(void* dst, const float* src, const int* counter, int dim, int maxSeqLen,
                                             cudaStream_t stream) {
    if (dim <= 0) return cudaSuccess;
    offset_memcpy_kernel<__half><<<zb::cdiv(dim, 256), 256, 0, stream>>>(static_cast<__half*>(dst), src, counter, dim, maxSeqLen);
    return cudaGetLastError();
}

ZB_API cudaError_t launch_increment_counter(int* counter, int delta, cudaStream_t stream) {
    counter_add_kernel<<<1, 1, 0, stream>>>(counter, delta);
    return cudaGetLastError();
}

ZB_API cudaError_t launch_reset_counter(int* counter, int value, cudaStream_t stream) {
    counter_set_kernel<<<1, 1, 0, stream>>>(counter, value);
    return cudaGetLastError();
}

// scratch: >= 2*ceil(n/256) 4-byte words (argmax.cu:90-93).  Stage 1 uses at
// most that many CTAs, capped at one wave of the chip.
ZB_API cudaError_t launch_argmax(const float* input, int* result, void* scratch, int n, cudaStream_t stream) {
    if (n <= 0) return cudaErrorInvalidValue;
    int cap = zb::cdiv(n, 256);
    int nb = cap < ZB_SMS * 2 ? cap : ZB_SMS * 2;
    float* bv = static_cast<float*>(scratch);
    int* bi = reinterpret_cast<int*>(static_cast<char*>(scratch) + sizeof(float) * cap);
    argmax_stage1<<<nb, 256, 0, stream>>>(input, bv, bi, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    argmax_stage2<<<1, 256, 0, stream>>>(bv, bi, result, nb);
    return cudaGetLastError();
}

ZB_API cudaError_t scaled_softmax_f32(const float* input, float* output, int outer, int inner, int axisSize,
                                      unsigned int scale_bits, cudaStream_t stream) {
    if (outer <= 0 || inner <= 0 || axisSize <= 0) return cudaSuccess;
    int threads = axisSize >= 256 ? 256 : (axisSize >= 128 ? 128 : (axisSize >= 64 ? 64 : 32));
    scaled_softmax_kernel<<<outer * inner, threads, 0, stream>>>(input, output, inner, axisSize, bits(scale_bits));
    return cudaGetLastError();
}

ZB_API cudaError_t launch_gather(const float* table, const long long* indices, float* output, int N, int D, int V,
                                 cudaStream_t stream) {
    if (N <= 0 || D <= 0) return cudaSuccess;
    gather_kernel<long long><<<N, D >= 256 ? 256 : 64, 0, stream>>>(table, indices, output, D, V);
    return cudaGetLastError();
}

ZB_API cudaError_t launch_gather_i32(const float* table, const int* indices, float* output, int N, int D, int V,
                                     cudaStream_t stream) {
    if (N <= 0 || D <= 0) return cudaSuccess;
    gather_kernel<int><<<N, D >= 256 ? 256 : 64, 0, stream>>>(table, indices, output, D, V);
    return cudaGetLastError();
}
