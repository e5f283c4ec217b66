// Boundary-only symbols: the reference's host loaders (internal/cuda/kernels/purego.go:141-273
// and ztensor's loader, see symbol_parity_test.go:16-29) dlsym every name below from the
// same libkernels.so and refuse to start if a REQUIRED one is missing, so a drop-in
// library must export them all.  They are off the decode hot path: elementwise/transposes
// are plain one-thread-per-element kernels (correct, not tuned); training-only and
// ONNX-only entry points report cudaErrorNotSupported instead of silently doing nothing.
//
// Reference definitions: internal/cuda/kernels/elementwise.cu, elementwise_fp16.cu,
// fp8_ops.cu, transpose.cu, dropout.cu, fused_adamw.cu, tiny_batched_gemm.cu,
// gemm_int4.cu, gemm_int8.cu.
#include <cuda_fp8.h>
#include <float.h>

#include "zb_common.cuh"

namespace {

inline float fbits(unsigned int b) {
    float f;
    memcpy(&f, &b, 4);
    return f;
}
inline int grid_for(int n) { return (n + 255) / 256; }

struct OpAdd { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpSub { __device__ float operator()(float a, float b) const { return a - b; } };
struct OpMul { __device__ float operator()(float a, float b) const { return a * b; } };
struct OpDiv { __device__ float operator()(float a, float b) const { return a / b; } };
struct OpPow { __device__ float operator()(float a, float b) const { return powf(a, b); } };

template <class Op>
__global__ void binary_kernel(const float* a, const float* b, float* c, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = Op()(a[i], b[i]);
}
template <class Op>
__global__ void scalar_kernel(const float* a, float s, float* c, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = Op()(a[i], s);
}
template <class Op>
__global__ void bcast2d_kernel(const float* a, const float* b, float* c, int sar, int sac, int sbr, int sbc, int M, int D) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * D) return;
    int m = i / D, d = i % D;
    c[i] = Op()(a[m * sar + d * sac], b[m * sbr + d * sbc]);
}
template <class Op>
__global__ void bcast4d_kernel(const float* a, const float* b, float* c, int d0, int d1, int d2, int d3, int sa0, int sa1, int sa2,
                               int sa3, int sb0, int sb1, int sb2, int sb3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d0 * d1 * d2 * d3) return;
    int i3 = i % d3, t = i / d3, i2 = t % d2;
    t /= d2;
    int i1 = t % d1, i0 = t / d1;
    c[i] = Op()(a[i0 * sa0 + i1 * sa1 + i2 * sa2 + i3 * sa3], b[i0 * sb0 + i1 * sb1 + i2 * sb2 + i3 * sb3]);
}

struct UExp { __device__ float operator()(float a) const { return expf(a); } };
struct ULog { __device__ float operator()(float a) const { return logf(a); } };
struct USqrt { __device__ float operator()(float a) const { return sqrtf(a); } };
struct URsqrt { __device__ float operator()(float a) const { return 1.0f / sqrtf(a); } };
struct USin { __device__ float operator()(float a) const { return sinf(a); } };
struct UCos { __device__ float operator()(float a) const { return cosf(a); } };
struct UTanh { __device__ float operator()(float a) const { return tanhf(a); } };
template <class Op>
__global__ void unary_kernel(const float* a, float* c, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = Op()(a[i]);
}
__global__ void tanh_prime_kernel(const float* a, const float* up, float* c, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float t = tanhf(a[i]);
        c[i] = (1.0f - t * t) * up[i];
    }
}
__global__ void fill_kernel(float* d, float v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = v;
}
__global__ void sum_axis_kernel(const float* in, float* out, int outer, int inner, int A) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= outer * inner) return;
    int o = i / inner, k = i % inner;
    float s = 0.0f;
    for (int a = 0; a < A; a++) s += in[((int64_t)o * A + a) * inner + k];
    out[i] = s;
}
__global__ void softmax_axis_kernel(const float* in, float* out, int outer, int inner, int A) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= outer * inner) return;
    int o = i / inner, k = i % inner;
    const float* x = in + (int64_t)o * A * inner + k;
    float* y = out + (int64_t)o * A * inner + k;
    float mx = -FLT_MAX;
    for (int a = 0; a < A; a++) mx = fmaxf(mx, x[(int64_t)a * inner]);
    float s = 0.0f;
    for (int a = 0; a < A; a++) {
        float e = expf(x[(int64_t)a * inner] - mx);
        y[(int64_t)a * inner] = e;
        s += e;
    }
    float inv = 1.0f / s;
    for (int a = 0; a < A; a++) y[(int64_t)a * inner] *= inv;
}
__global__ void repeat_kernel(const float* src, float* dst, int outer, int axis, int inner, int reps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= outer * axis * reps * inner) return;
    int in_i = i % inner, t = i / inner, ra = t % (axis * reps), o = t / (axis * reps);
    dst[i] = src[((int64_t)o * axis + ra / reps) * inner + in_i];
}

// ---- fp16 ------------------------------------------------------------------
template <class Op>
__global__ void binary_fp16_kernel(const __half* a, const __half* b, __half* c, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = __float2half(Op()(__half2float(a[i]), __half2float(b[i])));
}
__global__ void f32_to_fp16_kernel(const float* s, __half* d, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __float2half(s[i]);
}
__global__ void fp16_to_f32_kernel(const __half* s, float* d, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __half2float(s[i]);
}
template <typename In, typename W>
__global__ void rmsnorm_lowp_kernel(const In* in, const W* w, __half* out, float in_scale, float eps, int D) {
    __shared__ float red[32];
    int64_t row = blockIdx.x;
    float ss = 0.0f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        float v = float(in[row * D + i]) * in_scale;
        ss += v * v;
    }
    ss = zb::block_sum(ss, red);
    float s = rsqrtf(ss / (float)D + eps);
    for (int i = threadIdx.x; i < D; i += blockDim.x)
        out[row * D + i] = __float2half(float(in[row * D + i]) * in_scale * s * float(w[i]));
}
__global__ void scaled_softmax_fp16_kernel(const __half* in, __half* out, int inner, int A, float scale) {
    __shared__ float red[32];
    int o = blockIdx.x / inner, k = blockIdx.x % inner;
    const __half* x = in + (int64_t)o * A * inner + k;
    __half* y = out + (int64_t)o * A * inner + k;
    float mx = -FLT_MAX;
    for (int i = threadIdx.x; i < A; i += blockDim.x) mx = fmaxf(mx, __half2float(x[(int64_t)i * inner]) * scale);
    mx = zb::warp_max(mx);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0) red[w] = mx;
    __syncthreads();
    mx = zb::warp_max(l < nw ? red[l] : -FLT_MAX);
    float sum = 0.0f;
    for (int i = threadIdx.x; i < A; i += blockDim.x) sum += expf(__half2float(x[(int64_t)i * inner]) * scale - mx);
    sum = zb::block_sum(sum, red);
    float inv = 1.0f / sum;
    for (int i = threadIdx.x; i < A; i += blockDim.x)
        y[(int64_t)i * inner] = __float2half(expf(__half2float(x[(int64_t)i * inner]) * scale - mx) * inv);
}

// ---- fp8 e4m3 (fp8_ops.cu) ---------------------------------------------------
__device__ __forceinline__ float e4m3_to_f32(unsigned char v) {
    __nv_fp8_e4m3 t;
    t.__x = v;
    return float(t);
}
__global__ void fp8_dequant_kernel(const unsigned char* in, __half* out, float scale, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __hmul(__float2half(e4m3_to_f32(in[i])), __float2half(scale));
}
template <bool MUL>
__global__ void fp8_binary_kernel(const unsigned char* a, const unsigned char* b, __half* c, float sa, float sb, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __half va = __hmul(__float2half(e4m3_to_f32(a[i])), __float2half(sa));
    __half vb = __hmul(__float2half(e4m3_to_f32(b[i])), __float2half(sb));
    c[i] = MUL ? __hmul(va, vb) : __hadd(va, vb);
}
struct Fp8In {
    unsigned char v;
    __device__ operator float() const { return e4m3_to_f32(v); }
};

// ---- transposes ----------------------------------------------------------------
template <typename T>
__global__ void transpose2d_kernel(const T* in, T* out, int rows, int cols) {
    __shared__ T tile[32][33];
    int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8)
        if (x < cols && y0 + j < rows) tile[j][threadIdx.x] = in[(int64_t)(y0 + j) * cols + x];
    __syncthreads();
    int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
    for (int j = threadIdx.y; j < 32; j += 8)
        if (ox < rows && oy0 + j < cols) out[(int64_t)(oy0 + j) * rows + ox] = tile[threadIdx.x][j];
}
// out[flat] = in[sum_k coord_out[k] * in_strides[perm[k]]] (transpose.cu semantics:
// out_strides decompose the flat output index, perm maps output dims to input dims).
template <typename T>
__global__ void transpose_nd_kernel(const T* in, T* out, const int* in_strides, const int* out_strides, const int* perm, int ndim,
                                    int total) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int rem = i, src = 0;
    for (int k = 0; k < ndim; k++) {
        int c = rem / out_strides[k];
        rem -= c * out_strides[k];
        src += c * in_strides[perm[k]];
    }
    out[i] = in[src];
}

}  // namespace

#define ZB_BINARY(name, OP)                                                                                   \
    ZB_API cudaError_t name(const float* a, const float* b, float* c, int n, cudaStream_t stream) {          \
        if (n <= 0) return cudaSuccess;                                                                       \
        binary_kernel<OP><<<grid_for(n), 256, 0, stream>>>(a, b, c, n);                                       \
        return cudaGetLastError();                                                                            \
    }
ZB_BINARY(launch_add, OpAdd)
ZB_BINARY(launch_sub, OpSub)
ZB_BINARY(launch_mul, OpMul)
ZB_BINARY(launch_div, OpDiv)
ZB_BINARY(launch_pow, OpPow)

#define ZB_SCALAR(name, OP)                                                                                   \
    ZB_API cudaError_t name(const float* a, unsigned int scalar_bits, float* c, int n, cudaStream_t stream) { \
        if (n <= 0) return cudaSuccess;                                                                       \
        scalar_kernel<OP><<<grid_for(n), 256, 0, stream>>>(a, fbits(scalar_bits), c, n);                      \
        return cudaGetLastError();                                                                            \
    }
ZB_SCALAR(launch_add_scalar, OpAdd)
ZB_SCALAR(launch_mul_scalar, OpMul)
ZB_SCALAR(launch_div_scalar, OpDiv)
ZB_SCALAR(launch_sub_scalar, OpSub)
ZB_SCALAR(launch_pow_scalar, OpPow)

#define ZB_BCAST(name, OP)                                                                                               \
    ZB_API cudaError_t name(const float* a, const float* b, float* c, int stride_a_row, int stride_a_col, int stride_b_row, \
                            int stride_b_col, int M, int D, cudaStream_t stream) {                                      \
        if (M * D <= 0) return cudaSuccess;                                                                             \
        bcast2d_kernel<OP><<<grid_for(M * D), 256, 0, stream>>>(a, b, c, stride_a_row, stride_a_col, stride_b_row,       \
                                                                stride_b_col, M, D);                                    \
        return cudaGetLastError();                                                                                      \
    }
ZB_BCAST(launch_add_broadcast, OpAdd)
ZB_BCAST(launch_sub_broadcast, OpSub)
ZB_BCAST(launch_mul_broadcast, OpMul)
ZB_BCAST(launch_div_broadcast, OpDiv)

#define ZB_BCAST4(name, OP)                                                                                              \
    ZB_API cudaError_t name(const float* a, const float* b, float* c, int d0, int d1, int d2, int d3, int sa0, int sa1,   \
                            int sa2, int sa3, int sb0, int sb1, int sb2, int sb3, cudaStream_t stream) {                \
        int n = d0 * d1 * d2 * d3;                                                                                      \
        if (n <= 0) return cudaSuccess;                                                                                 \
        bcast4d_kernel<OP><<<grid_for(n), 256, 0, stream>>>(a, b, c, d0, d1, d2, d3, sa0, sa1, sa2, sa3, sb0, sb1, sb2, sb3); \
        return cudaGetLastError();                                                                                      \
    }
ZB_BCAST4(launch_add_broadcast4d, OpAdd)
ZB_BCAST4(launch_sub_broadcast4d, OpSub)
ZB_BCAST4(launch_mul_broadcast4d, OpMul)
ZB_BCAST4(launch_div_broadcast4d, OpDiv)

#define ZB_UNARY(name, OP)                                                                  \
    ZB_API cudaError_t name(const float* a, float* c, int n, cudaStream_t stream) {         \
        if (n <= 0) return cudaSuccess;                                                     \
        unary_kernel<OP><<<grid_for(n), 256, 0, stream>>>(a, c, n);                         \
        return cudaGetLastError();                                                          \
    }
ZB_UNARY(launch_exp, UExp)
ZB_UNARY(launch_log, ULog)
ZB_UNARY(launch_sqrt, USqrt)
ZB_UNARY(launch_rsqrt, URsqrt)
ZB_UNARY(launch_sin, USin)
ZB_UNARY(launch_cos, UCos)
ZB_UNARY(launch_tanh, UTanh)

ZB_API cudaError_t launch_tanh_prime(const float* a, const float* upstream, float* c, int n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    tanh_prime_kernel<<<grid_for(n), 256, 0, stream>>>(a, upstream, c, n);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_fill(float* data, unsigned int value_bits, int n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    fill_kernel<<<grid_for(n), 256, 0, stream>>>(data, fbits(value_bits), n);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_sum_axis(const float* input, float* output, int outer, int inner, int axisSize, cudaStream_t stream) {
    if (outer * inner <= 0) return cudaSuccess;
    sum_axis_kernel<<<grid_for(outer * inner), 256, 0, stream>>>(input, output, outer, inner, axisSize);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_softmax(const float* input, float* output, int outer, int inner, int axisSize, cudaStream_t stream) {
    if (outer * inner <= 0 || axisSize <= 0) return cudaSuccess;
    softmax_axis_kernel<<<grid_for(outer * inner), 256, 0, stream>>>(input, output, outer, inner, axisSize);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_repeat(const float* src, float* dst, int outerSize, int axisDim, int innerSize, int reps,
                                 cudaStream_t stream) {
    int n = outerSize * axisDim * reps * innerSize;
    if (n <= 0) return cudaSuccess;
    repeat_kernel<<<grid_for(n), 256, 0, stream>>>(src, dst, outerSize, axisDim, innerSize, reps);
    return cudaGetLastError();
}

#define ZB_BINARY16(name, OP)                                                                                                 \
    ZB_API cudaError_t name(const void* a, const void* b, void* c, int n, cudaStream_t stream) {                              \
        if (n <= 0) return cudaSuccess;                                                                                       \
        binary_fp16_kernel<OP><<<grid_for(n), 256, 0, stream>>>(static_cast<const __half*>(a), static_cast<const __half*>(b), \
                                                                static_cast<__half*>(c), n);                                  \
        return cudaGetLastError();                                                                                            \
    }
ZB_BINARY16(launch_add_fp16, OpAdd)
ZB_BINARY16(launch_sub_fp16, OpSub)
ZB_BINARY16(launch_mul_fp16, OpMul)
ZB_BINARY16(launch_div_fp16, OpDiv)

ZB_API cudaError_t launch_rmsnorm_fp16(const void* input, const void* weight, void* output, unsigned int eps_bits, int rows, int D,
                                       cudaStream_t stream) {
    if (rows <= 0 || D <= 0) return cudaSuccess;
    rmsnorm_lowp_kernel<__half, __half><<<rows, 256, 0, stream>>>(static_cast<const __half*>(input), static_cast<const __half*>(weight),
                                                                  static_cast<__half*>(output), 1.0f, fbits(eps_bits), D);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_scaled_softmax_fp16(const void* input, void* output, int outer, int inner, int axisSize,
                                              unsigned int scale_bits, cudaStream_t stream) {
    if (outer * inner <= 0 || axisSize <= 0) return cudaSuccess;
    scaled_softmax_fp16_kernel<<<outer * inner, 128, 0, stream>>>(static_cast<const __half*>(input), static_cast<__half*>(output), inner,
                                                                  axisSize, fbits(scale_bits));
    return cudaGetLastError();
}
ZB_API cudaError_t launch_f32_to_fp16(const void* src, void* dst, int n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    f32_to_fp16_kernel<<<grid_for(n), 256, 0, stream>>>(static_cast<const float*>(src), static_cast<__half*>(dst), n);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_fp16_to_f32(const void* src, void* dst, int n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    fp16_to_f32_kernel<<<grid_for(n), 256, 0, stream>>>(static_cast<const __half*>(src), static_cast<float*>(dst), n);
    return cudaGetLastError();
}

ZB_API cudaError_t launch_dequant_fp8e4m3_to_fp16(const void* input, void* output, unsigned int scale_bits, int n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    fp8_dequant_kernel<<<grid_for(n), 256, 0, stream>>>(static_cast<const unsigned char*>(input), static_cast<__half*>(output),
                                                        fbits(scale_bits), n);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_fp8_add(const void* a, const void* b, void* c, unsigned int scale_a_bits, unsigned int scale_b_bits, int n,
                                  cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    fp8_binary_kernel<false><<<grid_for(n), 256, 0, stream>>>(static_cast<const unsigned char*>(a), static_cast<const unsigned char*>(b),
                                                              static_cast<__half*>(c), fbits(scale_a_bits), fbits(scale_b_bits), n);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_fp8_mul(const void* a, const void* b, void* c, unsigned int scale_a_bits, unsigned int scale_b_bits, int n,
                                  cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    fp8_binary_kernel<true><<<grid_for(n), 256, 0, stream>>>(static_cast<const unsigned char*>(a), static_cast<const unsigned char*>(b),
                                                             static_cast<__half*>(c), fbits(scale_a_bits), fbits(scale_b_bits), n);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_fp8_rmsnorm(const void* input, const void* weight, void* output, unsigned int scale_bits,
                                      unsigned int eps_bits, int rows, int D, cudaStream_t stream) {
    if (rows <= 0 || D <= 0) return cudaSuccess;
    rmsnorm_lowp_kernel<Fp8In, __half><<<rows, 256, 0, stream>>>(static_cast<const Fp8In*>(input), static_cast<const __half*>(weight),
                                                                 static_cast<__half*>(output), fbits(scale_bits), fbits(eps_bits), D);
    return cudaGetLastError();
}

ZB_API cudaError_t launch_transpose_2d(const float* input, float* output, int rows, int cols, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    transpose2d_kernel<float><<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, stream>>>(input, output, rows, cols);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_transpose_2d_bf16(const unsigned short* input, unsigned short* output, int rows, int cols, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    transpose2d_kernel<unsigned short><<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, stream>>>(input, output, rows, cols);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_transpose_nd(const float* input, float* output, const int* in_strides, const int* out_strides, const int* perm,
                                       int ndim, int total, cudaStream_t stream) {
    if (total <= 0) return cudaSuccess;
    transpose_nd_kernel<float><<<grid_for(total), 256, 0, stream>>>(input, output, in_strides, out_strides, perm, ndim, total);
    return cudaGetLastError();
}
ZB_API cudaError_t launch_transpose_nd_bf16(const unsigned short* input, unsigned short* output, const int* in_strides,
                                            const int* out_strides, const int* perm, int ndim, int total, cudaStream_t stream) {
    if (total <= 0) return cudaSuccess;
    transpose_nd_kernel<unsigned short><<<grid_for(total), 256, 0, stream>>>(input, output, in_strides, out_strides, perm, ndim, total);
    return cudaGetLastError();
}

// ---- training-only / ONNX-only entry points: exported so dlsym succeeds, not implemented.
ZB_API cudaError_t dropout_f32(const float*, float*, int, uint32_t, uint64_t, int, uint32_t, cudaStream_t) { return cudaErrorNotSupported; }
ZB_API cudaError_t fused_adamw_f32(float*, float*, double*, float*, unsigned long long, unsigned long long, unsigned long long,
                                   unsigned long long, unsigned long long, unsigned long long, unsigned long long, int, cudaStream_t) {
    return cudaErrorNotSupported;
}
ZB_API cudaError_t tiny_batched_gemm_f32(const float*, const float*, float*, int, int, int, long long, long long, long long, int,
                                         cudaStream_t) {
    return cudaErrorNotSupported;
}
ZB_API cudaError_t gemm_int4_f32(const void*, const float*, float*, const float*, const void*, int, int, int, int, cudaStream_t) {
    return cudaErrorNotSupported;
}
ZB_API cudaError_t gemm_int4_f32_rmul(const void*, const float*, float*, const float*, const void*, int, int, int, int, cudaStream_t) {
    return cudaErrorNotSupported;
}
ZB_API cudaError_t gemm_int8_f32(const void*, const float*, float*, int, int, int, cudaStream_t) { return cudaErrorNotSupported; }
