// Batch-1 fused dequant-GEMV over GGUF block-quantized weights (sm_100a).
//
// y[M] = deq(W[M,K]) . x[K], f32 activations, f32 accumulation.  One warp
// owns ZB_ROWS rows at a time; lanes stride over 32-weight "chunks" of K so
// that consecutive lanes touch consecutive 16-byte pieces of the row (128-bit
// coalesced loads straight into registers), the activation chunk is read from
// shared memory once and reused for every row the warp owns, dequantisation
// happens in registers and the per-row partial sums meet in a warp shuffle.
// Grids are persistent: min(row groups, SMs x resident CTAs) CTAs loop over
// row groups, so x is staged once per CTA, not once per 16 rows.
//
// Replaces (same symbols, same argument order, reference file:line):
//   gemm_q4_f32      internal/cuda/kernels/gemm_q4.cu:163   (gemm_q4.h:32-36)
//   gemv_q4k_f32     internal/cuda/kernels/gemv_q4k.cu:146  (gemv_q4k.h:34-37)
//   gemv_q5k_f32     internal/cuda/kernels/gemv_q5k.cu:160
//   gemv_q6k_f32     internal/cuda/kernels/gemv_q6k.cu:135
//   gemm_q8_f32      internal/cuda/kernels/gemm_q8.cu:162
//   launch_sgemv_m1  internal/cuda/kernels/sgemv_m1.cu:89
//   dequant_q4k_f32  internal/cuda/kernels/dequant_q4k.cu:93
//   gemv_q4k_sm121_f32 / gemv_q4k_check_sm121  gemv_q4k_sm121.cu:196
#include "zb_common.cuh"
#include "zb_quant.cuh"

namespace {

using namespace zb;

constexpr int kWarps = 8;       // warps per CTA
constexpr int kRows = 2;        // rows per warp per pass (x chunk reused across them)
constexpr int kCtasPerSm = 4;   // persistent grid cap = ZB_SMS * kCtasPerSm

// ---------------------------------------------------------------------------
// Format traits.  A chunk is 32 weights that share their scale lookup; each
// trait says where the 32 matching activations live (load_x) and how to
// contract one row's chunk against them (dot).
// ---------------------------------------------------------------------------

struct XChunk {
    float v[32];
};

__device__ __forceinline__ void ld_x_run(float* dst, const float* sx, int k0, int n) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        if (i < n) {
            float4 t = *reinterpret_cast<const float4*>(sx + k0 + i);
            dst[i] = t.x; dst[i + 1] = t.y; dst[i + 2] = t.z; dst[i + 3] = t.w;
        }
    }
}

// 6-bit scale/min of sub-block j from the 12 packed bytes held as 3 words
// (gemv_q4k.cu:38-56 / gemv_q4k_test.go:18-27).
__device__ __forceinline__ void kq_scale_min(uint32_t s0, uint32_t s1, uint32_t s2, int j, float& sc, float& mn) {
    uint32_t a, b;
    if (j < 4) {
        a = (s0 >> (8 * j)) & 63u;
        b = (s1 >> (8 * j)) & 63u;
    } else {
        int jj = j - 4;
        a = ((s2 >> (8 * jj)) & 0xFu) | (((s0 >> (8 * jj + 6)) & 3u) << 4);
        b = ((s2 >> (8 * jj + 4)) & 0xFu) | (((s1 >> (8 * jj + 6)) & 3u) << 4);
    }
    sc = (float)a;
    mn = (float)b;
}

// ---- Q4_0, reference GPU "separated" layout (gemm_q4.h:3-4) ---------------
struct FmtQ4_0Sep {
    const uint16_t* scales;  // [M * K/32] fp16
    const uint8_t* data;     // [M * K/32 * 16]
    __device__ __forceinline__ void load_x(XChunk& xc, const float* sx, int chunk) const { ld_x_run(xc.v, sx, chunk * 32, 32); }
    __device__ __forceinline__ float dot(int64_t row, int chunks_per_row, int chunk, const XChunk& xc) const {
        int64_t blk = row * chunks_per_row + chunk;
        uint4 q = ldg128_stream(data + blk * 16);
        float d = h2f(ldg16(scales + blk));
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int lo = (int)((w[i] >> (8 * b)) & 0xFu) - 8;
                int hi = (int)((w[i] >> (8 * b + 4)) & 0xFu) - 8;
                s = fmaf((float)lo, xc.v[4 * i + b], s);
                s = fmaf((float)hi, xc.v[16 + 4 * i + b], s);
            }
        }
        return s * d;
    }
};

// ---- Q4_K raw GGUF super-blocks, 144 B / 256 (gemv_q4k.cu:8-21) -----------
struct FmtQ4_K {
    const uint8_t* w;
    // chunk c of block b: g = c>>1, half = c&1 -> 16 low-nibble weights at
    // k = 64g + 16half + i (sub-block 2g) and 16 high-nibble at +32 (2g+1).
    __device__ __forceinline__ void load_x(XChunk& xc, const float* sx, int chunk) const {
        int b = chunk >> 3, c = chunk & 7;
        int k0 = b * 256 + (c >> 1) * 64 + (c & 1) * 16;
        ld_x_run(xc.v, sx, k0, 16);
        ld_x_run(xc.v + 16, sx, k0 + 32, 16);
    }
    __device__ __forceinline__ float dot(int64_t row, int chunks_per_row, int chunk, const XChunk& xc) const {
        int b = chunk >> 3, c = chunk & 7, g = c >> 1;
        const uint8_t* blk = w + (row * (chunks_per_row >> 3) + b) * 144;
        uint4 hdr = ldg128(blk);
        uint4 q = ldg128_stream(blk + 16 + c * 16);
        float d = __half2float(__ushort_as_half((uint16_t)(hdr.x & 0xFFFF)));
        float dmin = __half2float(__ushort_as_half((uint16_t)(hdr.x >> 16)));
        float sc0, mn0, sc1, mn1;
        kq_scale_min(hdr.y, hdr.z, hdr.w, 2 * g, sc0, mn0);
        kq_scale_min(hdr.y, hdr.z, hdr.w, 2 * g + 1, sc1, mn1);
        uint32_t wd[4] = {q.x, q.y, q.z, q.w};
        float slo = 0.0f, shi = 0.0f, xlo = 0.0f, xhi = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                float lo = (float)((wd[i] >> (8 * t)) & 0xFu);
                float hi = (float)((wd[i] >> (8 * t + 4)) & 0xFu);
                slo = fmaf(lo, xc.v[4 * i + t], slo);
                shi = fmaf(hi, xc.v[16 + 4 * i + t], shi);
                xlo += xc.v[4 * i + t];
                xhi += xc.v[16 + 4 * i + t];
            }
        }
        // sum (d*sc*q - dmin*m) x = d*sc*sum(q x) - dmin*m*sum(x); d*sc and dmin*m are exact in f32.
        return (d * sc0) * slo - (dmin * mn0) * xlo + (d * sc1) * shi - (dmin * mn1) * xhi;
    }
};

// ---- Q5_K raw GGUF super-blocks, 176 B / 256 (gemv_q5k.cu:7-23) -----------
struct FmtQ5_K {
    const uint8_t* w;
    __device__ __forceinline__ void load_x(XChunk& xc, const float* sx, int chunk) const {
        int b = chunk >> 3, c = chunk & 7;
        int k0 = b * 256 + (c >> 1) * 64 + (c & 1) * 16;
        ld_x_run(xc.v, sx, k0, 16);
        ld_x_run(xc.v + 16, sx, k0 + 32, 16);
    }
    __device__ __forceinline__ float dot(int64_t row, int chunks_per_row, int chunk, const XChunk& xc) const {
        int b = chunk >> 3, c = chunk & 7, g = c >> 1, half = c & 1;
        const uint8_t* blk = w + (row * (chunks_per_row >> 3) + b) * 176;
        uint4 hdr = ldg128(blk);
        uint4 q = ldg128_stream(blk + 16 + c * 16);
        uint4 h = ldg128(blk + 144 + half * 16);
        float d = __half2float(__ushort_as_half((uint16_t)(hdr.x & 0xFFFF)));
        float dmin = __half2float(__ushort_as_half((uint16_t)(hdr.x >> 16)));
        float sc0, mn0, sc1, mn1;
        kq_scale_min(hdr.y, hdr.z, hdr.w, 2 * g, sc0, mn0);
        kq_scale_min(hdr.y, hdr.z, hdr.w, 2 * g + 1, sc1, mn1);
        uint32_t wd[4] = {q.x, q.y, q.z, q.w};
        uint32_t hb[4] = {h.x >> (2 * g), h.y >> (2 * g), h.z >> (2 * g), h.w >> (2 * g)};
        float slo = 0.0f, shi = 0.0f, xlo = 0.0f, xhi = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                uint32_t lo = ((wd[i] >> (8 * t)) & 0xFu) | (((hb[i] >> (8 * t)) & 1u) << 4);
                uint32_t hi = ((wd[i] >> (8 * t + 4)) & 0xFu) | (((hb[i] >> (8 * t + 1)) & 1u) << 4);
                slo = fmaf((float)lo, xc.v[4 * i + t], slo);
                shi = fmaf((float)hi, xc.v[16 + 4 * i + t], shi);
                xlo += xc.v[4 * i + t];
                xhi += xc.v[16 + 4 * i + t];
            }
        }
        return (d * sc0) * slo - (dmin * mn0) * xlo + (d * sc1) * shi - (dmin * mn1) * xhi;
    }
};

// ---- Q6_K raw GGUF super-blocks, 210 B / 256 (gemv_q6k.cu:7-25) -----------
// Rows are only 2-byte aligned, so the raw-layout path uses 16-bit loads.
struct FmtQ6_K {
    const uint8_t* w;
    // chunk c: half = c>>2, l0 = 8*(c&3): weights (half, quarter q, l0..l0+7), q = 0..3
    __device__ __forceinline__ void load_x(XChunk& xc, const float* sx, int chunk) const {
        int b = chunk >> 3, c = chunk & 7;
        int k0 = b * 256 + (c >> 2) * 128 + (c & 3) * 8;
#pragma unroll
        for (int q = 0; q < 4; q++) ld_x_run(xc.v + 8 * q, sx, k0 + 32 * q, 8);
    }
    __device__ __forceinline__ float dot(int64_t row, int chunks_per_row, int chunk, const XChunk& xc) const {
        int b = chunk >> 3, c = chunk & 7, half = c >> 2, l0 = (c & 3) * 8;
        const uint8_t* blk = w + (row * (chunks_per_row >> 3) + b) * 210;
        const uint16_t* pa = reinterpret_cast<const uint16_t*>(blk + half * 64 + l0);
        const uint16_t* pb = reinterpret_cast<const uint16_t*>(blk + half * 64 + 32 + l0);
        const uint16_t* ph = reinterpret_cast<const uint16_t*>(blk + 128 + half * 32 + l0);
        uint32_t A[2], B[2], H[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            A[i] = (uint32_t)__ldg(pa + 2 * i) | ((uint32_t)__ldg(pa + 2 * i + 1) << 16);
            B[i] = (uint32_t)__ldg(pb + 2 * i) | ((uint32_t)__ldg(pb + 2 * i + 1) << 16);
            H[i] = (uint32_t)__ldg(ph + 2 * i) | ((uint32_t)__ldg(ph + 2 * i + 1) << 16);
        }
        const int8_t* sc = reinterpret_cast<const int8_t*>(blk + 192) + half * 8 + ((c & 3) >> 1);
        float d = h2f(ldg16(blk + 208));
        float s0 = d * (float)__ldg(sc), s2 = d * (float)__ldg(sc + 2), s4 = d * (float)__ldg(sc + 4), s6 = d * (float)__ldg(sc + 6);
        float a1 = 0.0f, a2 = 0.0f, a3 = 0.0f, a4 = 0.0f;
#pragma unroll
        for (int i = 0; i < 2; i++) {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                uint32_t av = (A[i] >> (8 * t)) & 0xFFu, bv = (B[i] >> (8 * t)) & 0xFFu, hv = (H[i] >> (8 * t)) & 0xFFu;
                int q1 = (int)((av & 0xFu) | ((hv & 3u) << 4)) - 32;
                int q2 = (int)((bv & 0xFu) | (((hv >> 2) & 3u) << 4)) - 32;
                int q3 = (int)((av >> 4) | (((hv >> 4) & 3u) << 4)) - 32;
                int q4 = (int)((bv >> 4) | (((hv >> 6) & 3u) << 4)) - 32;
                int l = 4 * i + t;
                a1 = fmaf((float)q1, xc.v[l], a1);
                a2 = fmaf((float)q2, xc.v[8 + l], a2);
                a3 = fmaf((float)q3, xc.v[16 + l], a3);
                a4 = fmaf((float)q4, xc.v[24 + l], a4);
            }
        }
        return s0 * a1 + s2 * a2 + s4 * a3 + s6 * a4;
    }
};

// ---- Q8_0, reference device layout: f32 scale + 32 int8 = 36 B (gemm_q8.cu:1-7)
struct FmtQ8_36 {
    const uint8_t* w;
    __device__ __forceinline__ void load_x(XChunk& xc, const float* sx, int chunk) const { ld_x_run(xc.v, sx, chunk * 32, 32); }
    __device__ __forceinline__ float dot(int64_t row, int chunks_per_row, int chunk, const XChunk& xc) const {
        const uint32_t* blk = reinterpret_cast<const uint32_t*>(w + (row * chunks_per_row + chunk) * 36);
        float d = __uint_as_float(__ldg(blk));
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t v = ldg32_stream(blk + 1 + i);
#pragma unroll
            for (int t = 0; t < 4; t++) {
                int q = (int)(int8_t)((v >> (8 * t)) & 0xFFu);
                s = fmaf((float)q, xc.v[4 * i + t], s);
            }
        }
        return s * d;
    }
};

// ---- Q4_0 / Q8_0 in their on-disk GGUF layouts (18 B / 34 B blocks, 2-byte
// aligned): the zb_* entry points that skip the reference's upload-time repack.
struct FmtQ4_0Raw {
    const uint8_t* w;
    __device__ __forceinline__ void load_x(XChunk& xc, const float* sx, int chunk) const { ld_x_run(xc.v, sx, chunk * 32, 32); }
    __device__ __forceinline__ float dot(int64_t row, int chunks_per_row, int chunk, const XChunk& xc) const {
        const uint16_t* blk = reinterpret_cast<const uint16_t*>(w + (row * chunks_per_row + chunk) * 18);
        float d = h2f(__ldg(blk));
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t v = __ldg(blk + 1 + i);
#pragma unroll
            for (int t = 0; t < 2; t++) {
                uint32_t byte = (v >> (8 * t)) & 0xFFu;
                s = fmaf((float)((int)(byte & 0xFu) - 8), xc.v[2 * i + t], s);
                s = fmaf((float)((int)(byte >> 4) - 8), xc.v[16 + 2 * i + t], s);
            }
        }
        return s * d;
    }
};
struct FmtQ8_0Raw {
    const uint8_t* w;
    __device__ __forceinline__ void load_x(XChunk& xc, const float* sx, int chunk) const { ld_x_run(xc.v, sx, chunk * 32, 32); }
    __device__ __forceinline__ float dot(int64_t row, int chunks_per_row, int chunk, const XChunk& xc) const {
        const uint16_t* blk = reinterpret_cast<const uint16_t*>(w + (row * chunks_per_row + chunk) * 34);
        float d = h2f(__ldg(blk));
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            uint32_t v = __ldg(blk + 1 + i);
            s = fmaf((float)(int)(int8_t)(v & 0xFFu), xc.v[2 * i], s);
            s = fmaf((float)(int)(int8_t)(v >> 8), xc.v[2 * i + 1], s);
        }
        return s * d;
    }
};

// ---- plain f32 rows (launch_sgemv_m1; router / un-quantized matrices) -----
struct FmtF32 {
    const float* w;
    int k;
    __device__ __forceinline__ void load_x(XChunk& xc, const float* sx, int chunk) const { ld_x_run(xc.v, sx, chunk * 32, 32); }
    __device__ __forceinline__ float dot(int64_t row, int chunks_per_row, int chunk, const XChunk& xc) const {
        const float* p = w + row * (int64_t)k + chunk * 32;
        float s = 0.0f;
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint4 u = ldg128_stream(p + 4 * i);
                s = fmaf(__uint_as_float(u.x), xc.v[4 * i], s);
                s = fmaf(__uint_as_float(u.y), xc.v[4 * i + 1], s);
                s = fmaf(__uint_as_float(u.z), xc.v[4 * i + 2], s);
                s = fmaf(__uint_as_float(u.w), xc.v[4 * i + 3], s);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) s = fmaf(__ldg(p + i), xc.v[i], s);
        }
        return s;
    }
};

template <class F>
__global__ void __launch_bounds__(kWarps * 32) gemv_kernel(F f, const float* __restrict__ x, float* __restrict__ y, int M, int K) {
    extern __shared__ __align__(16) float sx[];
    for (int i = threadIdx.x; i < K; i += blockDim.x) sx[i] = x[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = K / 32;
    const int groups = (M + kRows - 1) / kRows;
    for (int grp = blockIdx.x * kWarps + warp; grp < groups; grp += gridDim.x * kWarps) {
        int r0 = grp * kRows;
        float acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; r++) acc[r] = 0.0f;
        for (int c = lane; c < chunks; c += 32) {
            XChunk xc;
            f.load_x(xc, sx, c);
#pragma unroll
            for (int r = 0; r < kRows; r++)
                if (r0 + r < M) acc[r] += f.dot(r0 + r, chunks, c, xc);
        }
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            float s = warp_sum(acc[r]);
            if (lane == 0 && r0 + r < M) y[r0 + r] = s;
        }
    }
}

// K that is a multiple of 32 but whose tail does not fill 32 (F32 only) is
// handled by the caller padding check; quantized formats require K % block == 0.
template <class F>
cudaError_t launch_gemv(const F& f, const float* x, float* y, int M, int K, cudaStream_t stream) {
    if (M <= 0 || K <= 0) return cudaSuccess;
    size_t smem = (size_t)K * sizeof(float);
    static DeviceOnce once;  // per instantiation and per device; attribute calls are capture-safe
    if (smem > 48 * 1024)
        if (cudaError_t e = once.ensure(smem, [&] { return cudaFuncSetAttribute(gemv_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }))
            return e;
    int groups = cdiv(M, kRows);
    int grid = cdiv(groups, kWarps);
    int cap = ZB_SMS * kCtasPerSm;
    if (grid > cap) grid = cap;
    gemv_kernel<F><<<grid, kWarps * 32, smem, stream>>>(f, x, y, M, K);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// N > 1 fallbacks of the two reference "gemm" entry points: C[M,N] =
// deq(A[M,K]) . B[K,N] with B row-major.  The B200 batched path is the
// tcgen05 kernel in gemm_tc.cu; these keep the drop-in symbols complete.
// ---------------------------------------------------------------------------
__global__ void gemm_q4_sep_kernel(const uint16_t* scales, const uint8_t* data, const float* __restrict__ B,
                                   float* __restrict__ C, int M, int K, int N) {
    int n = blockIdx.x * 32 + (threadIdx.x & 31);
    int m = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (m >= M) return;
    int nblk = K / 32;
    float acc = 0.0f;
    for (int b = 0; b < nblk; b++) {
        int64_t blk = (int64_t)m * nblk + b;
        float d = h2f(ldg16(scales + blk));
        uint4 q = ldg128(data + blk * 16);
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
        float s = 0.0f;
        if (n < N) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                uint32_t byte = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                s = fmaf((float)((int)(byte & 0xFu) - 8), B[(int64_t)(b * 32 + i) * N + n], s);
                s = fmaf((float)((int)(byte >> 4) - 8), B[(int64_t)(b * 32 + 16 + i) * N + n], s);
            }
        }
        acc = fmaf(s, d, acc);
    }
    if (n < N) C[(int64_t)m * N + n] = acc;
}

__global__ void gemm_q8_36_kernel(const uint8_t* A, const float* __restrict__ B, float* __restrict__ C, int M, int K, int N) {
    int n = blockIdx.x * 32 + (threadIdx.x & 31);
    int m = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (m >= M) return;
    int nblk = K / 32;
    float acc = 0.0f;
    for (int b = 0; b < nblk; b++) {
        const uint32_t* blk = reinterpret_cast<const uint32_t*>(A + ((int64_t)m * nblk + b) * 36);
        float d = __uint_as_float(__ldg(blk));
        float s = 0.0f;
        if (n < N) {
#pragma unroll
            for (int i = 0; i < 32; i++) {
                int q = (int)(int8_t)((__ldg(blk + 1 + (i >> 2)) >> (8 * (i & 3))) & 0xFFu);
                s = fmaf((float)q, B[(int64_t)(b * 32 + i) * N + n], s);
            }
        }
        acc = fmaf(s, d, acc);
    }
    if (n < N) C[(int64_t)m * N + n] = acc;
}

// Q4_K -> f32, one CTA per super-block (dequant_q4k.cu:38-91 semantics).
__global__ void dequant_q4k_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t nblk) {
    int64_t b = blockIdx.x;
    if (b >= nblk) return;
    const uint8_t* blk = src + b * 144;
    int t = threadIdx.x;          // 0..255 -> element index
    int g = t >> 6, within = t & 63, l = within & 31, hi = within >> 5;
    uint32_t s0 = __ldg(reinterpret_cast<const uint32_t*>(blk + 4));
    uint32_t s1 = __ldg(reinterpret_cast<const uint32_t*>(blk + 8));
    uint32_t s2 = __ldg(reinterpret_cast<const uint32_t*>(blk + 12));
    float d = h2f(ldg16(blk)), dmin = h2f(ldg16(blk + 2));
    float sc, mn;
    kq_scale_min(s0, s1, s2, 2 * g + hi, sc, mn);
    uint8_t q = __ldg(blk + 16 + g * 32 + l);
    float v = (float)(hi ? (q >> 4) : (q & 0xF));
    dst[b * 256 + t] = __fsub_rn(__fmul_rn(__fmul_rn(d, sc), v), __fmul_rn(dmin, mn));
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================

ZB_API cudaError_t gemm_q4_f32(const void* A_q4, const float* B, float* C, int M, int K, int N, int data_offset,
                               cudaStream_t stream) {
    if (K % 32) return cudaErrorInvalidValue;
    const uint16_t* scales = static_cast<const uint16_t*>(A_q4);
    const uint8_t* data = static_cast<const uint8_t*>(A_q4) + data_offset;
    if (N == 1) {
        FmtQ4_0Sep f{scales, data};
        return launch_gemv(f, B, C, M, K, stream);
    }
    dim3 grid(zb::cdiv(N, 32), zb::cdiv(M, 8));
    gemm_q4_sep_kernel<<<grid, 256, 0, stream>>>(scales, data, B, C, M, K, N);
    return cudaGetLastError();
}

ZB_API cudaError_t gemv_q4k_f32(const void* W, const float* x, float* y, int M, int K, cudaStream_t stream) {
    if (K % 256) return cudaErrorInvalidValue;
    FmtQ4_K f{static_cast<const uint8_t*>(W)};
    return launch_gemv(f, x, y, M, K, stream);
}

// The reference ships a GB10-only variant behind these two names
// (gemv_q4k_sm121.cu:196); on B200 they resolve to the sm_100a kernel.
ZB_API cudaError_t gemv_q4k_sm121_f32(const void* W, const float* x, float* y, int M, int K, cudaStream_t stream) {
    return gemv_q4k_f32(W, x, y, M, K, stream);
}
ZB_API int gemv_q4k_check_sm121() { return 0; }

ZB_API cudaError_t gemv_q5k_f32(const void* W, const float* x, float* y, int M, int K, cudaStream_t stream) {
    if (K % 256) return cudaErrorInvalidValue;
    FmtQ5_K f{static_cast<const uint8_t*>(W)};
    return launch_gemv(f, x, y, M, K, stream);
}

ZB_API cudaError_t gemv_q6k_f32(const void* W, const float* x, float* y, int M, int K, cudaStream_t stream) {
    if (K % 256) return cudaErrorInvalidValue;
    FmtQ6_K f{static_cast<const uint8_t*>(W)};
    return launch_gemv(f, x, y, M, K, stream);
}

ZB_API cudaError_t gemm_q8_f32(const void* A_q8, const float* B, float* C, int M, int K, int N, cudaStream_t stream) {
    if (K % 32) return cudaErrorInvalidValue;
    if (N == 1) {
        FmtQ8_36 f{static_cast<const uint8_t*>(A_q8)};
        return launch_gemv(f, B, C, M, K, stream);
    }
    dim3 grid(zb::cdiv(N, 32), zb::cdiv(M, 8));
    gemm_q8_36_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(A_q8), B, C, M, K, N);
    return cudaGetLastError();
}

namespace {
__global__ void sgemv_tail_kernel(float* y, const float* A, const float* x, int M, int N) {
    // generic N (not a multiple of 32): one warp per row
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    float s = 0.0f;
    for (int i = lane; i < N; i += 32) s = fmaf(__ldg(A + (int64_t)row * N + i), __ldg(x + i), s);
    s = zb::warp_sum(s);
    if (lane == 0) y[row] = s;
}
}  // namespace

// y[M] = A[M,N] . x[N]  (sgemv_m1.cu:89; note the reference names the inner dim N).
ZB_API cudaError_t launch_sgemv_m1(float* y, const float* A, const float* x, int M, int N, cudaStream_t stream) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    if (N % 32 == 0) {
        FmtF32 f{A, N};
        return launch_gemv(f, x, y, M, N, stream);
    }
    sgemv_tail_kernel<<<zb::cdiv(M, 8), 256, 0, stream>>>(y, A, x, M, N);
    return cudaGetLastError();
}

ZB_API cudaError_t dequant_q4k_f32(const void* src, float* dst, int rows, int K, cudaStream_t stream) {
    if (K % 256) return cudaErrorInvalidValue;
    int64_t nblk = (int64_t)rows * (K / 256);
    if (nblk == 0) return cudaSuccess;
    dequant_q4k_kernel<<<(unsigned)nblk, 256, 0, stream>>>(static_cast<const uint8_t*>(src), dst, nblk);
    return cudaGetLastError();
}

// ---- zb_* entry points (include/zb200.h) ------------------------------------
ZB_API int zb_gemv_q4_0_f32(const void* W, const float* x, float* y, int M, int K, cudaStream_t stream) {
    if (K % 32) return cudaErrorInvalidValue;
    FmtQ4_0Raw f{static_cast<const uint8_t*>(W)};
    return launch_gemv(f, x, y, M, K, stream);
}
ZB_API int zb_gemv_q8_0_f32(const void* W, const float* x, float* y, int M, int K, cudaStream_t stream) {
    if (K % 32) return cudaErrorInvalidValue;
    FmtQ8_0Raw f{static_cast<const uint8_t*>(W)};
    return launch_gemv(f, x, y, M, K, stream);
}

namespace {
__global__ void dequant_any_kernel(int qtype, const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = zb::deq_raw(qtype, src, i);
}
}  // namespace

ZB_API int zb_dequant_f32(int qtype, const void* src, float* dst, int64_t n, cudaStream_t stream) {
    int be = zb::block_elems(qtype);
    if (be == 0 || n % be) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    int grid = (int)(blocks < ZB_SMS * 16 ? blocks : ZB_SMS * 16);
    dequant_any_kernel<<<grid, 256, 0, stream>>>(qtype, static_cast<const uint8_t*>(src), dst, n);
    return cudaGetLastError();
}
