// Attention kernels (sm_100a): split-KV flash decode over the GPU KV cache,
// per-head decode, and causal prefill.  f32 in / f32 accumulate, online
// softmax (running max + rescale), log-sum-exp merge across warps and splits.
//
// Replaces (reference file:line):
//   flash_decode_splitkv_f32     internal/cuda/kernels/flash_decode.cu:240 (flash_decode.h:39-46)
//   flash_attention_decode_f32   internal/cuda/kernels/flash_attention.cu:344
//   flash_attention_forward_f32  internal/cuda/kernels/flash_attention.cu:155
// and the production decode path MatMulTransposeB + GPUFusedSoftmaxVMul after a
// K/V Repeat (layers/attention/grouped_query_attention.go:1017-1046): here GQA
// query heads index their KV head directly, nothing is replicated in HBM.
#include <float.h>

#include "zb_common.cuh"

namespace {

using namespace zb;

constexpr int kDecWarps = 4;
constexpr int kMaxHd = 256;
constexpr int kPerLane = kMaxHd / 32;  // 8 floats per lane at hd = 256

struct Online {
    float m, l;
    float acc[kPerLane];
};

__device__ __forceinline__ void online_init(Online& o) {
    o.m = -FLT_MAX;
    o.l = 0.0f;
#pragma unroll
    for (int i = 0; i < kPerLane; i++) o.acc[i] = 0.0f;
}

// Lane owns head-dim elements {lane*4 + 128*j + (0..3)}: 128-bit coalesced K/V reads.
template <int NV>  // float4 vectors per lane: hd = 128*NV (NV = 1, 2); NV = 0 -> generic strided
__device__ __forceinline__ void load_vec(float* dst, const float* base, int hd, int lane) {
    if (NV > 0) {
#pragma unroll
        for (int j = 0; j < NV; j++) {
            float4 t = *reinterpret_cast<const float4*>(base + j * 128 + lane * 4);
            dst[4 * j] = t.x; dst[4 * j + 1] = t.y; dst[4 * j + 2] = t.z; dst[4 * j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kPerLane; j++) {
            int d = lane + 32 * j;
            dst[j] = d < hd ? base[d] : 0.0f;
        }
    }
}
template <int NV>
__device__ __forceinline__ void store_vec(float* base, const float* src, int hd, int lane) {
    if (NV > 0) {
#pragma unroll
        for (int j = 0; j < NV; j++)
            *reinterpret_cast<float4*>(base + j * 128 + lane * 4) = make_float4(src[4 * j], src[4 * j + 1], src[4 * j + 2], src[4 * j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < kPerLane; j++) {
            int d = lane + 32 * j;
            if (d < hd) base[d] = src[j];
        }
    }
}
template <int NV>
constexpr int nvals() { return NV > 0 ? 4 * NV : kPerLane; }

// One warp walks positions [t0, t1) with stride `step`, one position at a time:
// s = (q.k)*scale, online-softmax update, acc += p*v.
template <int NV>
__device__ __forceinline__ void walk(Online& o, const float* q, const float* K, const float* V, int64_t stride, int t0, int t1,
                                     int step, int hd, int lane, float scale) {
    constexpr int N = nvals<NV>();
    for (int t = t0; t < t1; t += step) {
        float kv[N];
        load_vec<NV>(kv, K + (int64_t)t * stride, hd, lane);
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < N; i++) s = fmaf(q[i], kv[i], s);
        s = warp_sum(s) * scale;
        float mn = fmaxf(o.m, s);
        float corr = __expf(o.m - mn), p = __expf(s - mn);
        load_vec<NV>(kv, V + (int64_t)t * stride, hd, lane);
        o.l = o.l * corr + p;
#pragma unroll
        for (int i = 0; i < N; i++) o.acc[i] = fmaf(p, kv[i], o.acc[i] * corr);
        o.m = mn;
    }
}

// Merge the warps of a CTA through shared memory; result lands in warp 0.
template <int NV>
__device__ __forceinline__ void merge_warps(Online& o, float* sm_m, float* sm_l, float* sm_acc, int hd, int warp, int lane, int nwarps) {
    constexpr int N = nvals<NV>();
    if (lane == 0) { sm_m[warp] = o.m; sm_l[warp] = o.l; }
    store_vec<NV>(sm_acc + warp * kMaxHd, o.acc, hd, lane);
    __syncthreads();
    if (warp == 0) {
        float m = -FLT_MAX;
        for (int w = 0; w < nwarps; w++) m = fmaxf(m, sm_m[w]);
        float l = 0.0f, acc[N];
#pragma unroll
        for (int i = 0; i < N; i++) acc[i] = 0.0f;
        for (int w = 0; w < nwarps; w++) {
            float c = sm_l[w] > 0.0f ? __expf(sm_m[w] - m) : 0.0f;
            l += sm_l[w] * c;
            float v[N];
            load_vec<NV>(v, sm_acc + w * kMaxHd, hd, lane);
#pragma unroll
            for (int i = 0; i < N; i++) acc[i] = fmaf(v[i], c, acc[i]);
        }
        o.m = m; o.l = l;
#pragma unroll
        for (int i = 0; i < N; i++) o.acc[i] = acc[i];
    }
}

// ---- split-KV decode, cache layout [batch, max_kv, nKV*hd] ----------------
template <int NV>
__global__ void __launch_bounds__(kDecWarps * 32) decode_split_kernel(
    const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, float* __restrict__ pO,
    float* __restrict__ pM, float* __restrict__ pL, int max_kv, int hd, int kv_len, const int* __restrict__ kv_len_ptr, int nQ,
    int nKV, int chunk, int splits) {
    __shared__ float sm_m[kDecWarps], sm_l[kDecWarps];
    __shared__ __align__(16) float sm_acc[kDecWarps * kMaxHd];
    constexpr int N = nvals<NV>();
    int bh = blockIdx.x, split = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int len = kv_len_ptr ? *kv_len_ptr : kv_len;
    if (len > max_kv) len = max_kv;
    int b = bh / nQ, h = bh % nQ, kvh = h / (nQ / nKV);
    int64_t stride = (int64_t)nKV * hd;
    const float* Kb = K + (int64_t)b * max_kv * stride + (int64_t)kvh * hd;
    const float* Vb = V + (int64_t)b * max_kv * stride + (int64_t)kvh * hd;
    float q[N];
    load_vec<NV>(q, Q + (int64_t)bh * hd, hd, lane);
    int t0 = split * chunk, t1 = min(t0 + chunk, len);
    Online o;
    online_init(o);
    walk<NV>(o, q, Kb, Vb, stride, t0 + warp, t1, kDecWarps, hd, lane, rsqrtf((float)hd));
    merge_warps<NV>(o, sm_m, sm_l, sm_acc, hd, warp, lane, kDecWarps);
    if (warp == 0) {
        int64_t slot = (int64_t)bh * splits + split;
        store_vec<NV>(pO + slot * hd, o.acc, hd, lane);
        if (lane == 0) { pM[slot] = o.m; pL[slot] = o.l; }
    }
}

__global__ void decode_reduce_kernel(const float* __restrict__ pO, const float* __restrict__ pM, const float* __restrict__ pL,
                                     float* __restrict__ O, int hd, int splits) {
    int bh = blockIdx.x;
    float m = -FLT_MAX;
    for (int s = 0; s < splits; s++)
        if (pL[(int64_t)bh * splits + s] > 0.0f) m = fmaxf(m, pM[(int64_t)bh * splits + s]);
    float l = 0.0f;
    for (int s = 0; s < splits; s++) {
        float ls = pL[(int64_t)bh * splits + s];
        if (ls > 0.0f) l += ls * __expf(pM[(int64_t)bh * splits + s] - m);
    }
    float inv = l > 0.0f ? 1.0f / l : 0.0f;
    for (int d = threadIdx.x; d < hd; d += blockDim.x) {
        float acc = 0.0f;
        for (int s = 0; s < splits; s++) {
            int64_t slot = (int64_t)bh * splits + s;
            float ls = pL[slot];
            if (ls > 0.0f) acc = fmaf(pO[slot * hd + d], __expf(pM[slot] - m), acc);
        }
        O[(int64_t)bh * hd + d] = acc * inv;
    }
}

// ---- per-head decode, cache layout [batch*nKV, max_kv, hd] ----------------
template <int NV>
__global__ void __launch_bounds__(kDecWarps * 32) decode_head_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                                     const float* __restrict__ V, float* __restrict__ O, int max_kv,
                                                                     int hd, int kv_len, const int* __restrict__ kv_len_ptr, int nQ,
                                                                     int nKV) {
    __shared__ float sm_m[kDecWarps], sm_l[kDecWarps];
    __shared__ __align__(16) float sm_acc[kDecWarps * kMaxHd];
    constexpr int N = nvals<NV>();
    int bh = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int len = kv_len_ptr ? *kv_len_ptr : kv_len;
    if (len > max_kv) len = max_kv;
    int b = bh / nQ, h = bh % nQ, kvh = h / (nQ / nKV);
    const float* Kb = K + ((int64_t)b * nKV + kvh) * max_kv * hd;
    const float* Vb = V + ((int64_t)b * nKV + kvh) * max_kv * hd;
    float q[N];
    load_vec<NV>(q, Q + (int64_t)bh * hd, hd, lane);
    Online o;
    online_init(o);
    walk<NV>(o, q, Kb, Vb, hd, warp, len, kDecWarps, hd, lane, rsqrtf((float)hd));
    merge_warps<NV>(o, sm_m, sm_l, sm_acc, hd, warp, lane, kDecWarps);
    if (warp == 0) {
        float inv = o.l > 0.0f ? 1.0f / o.l : 0.0f;
        float out[N];
#pragma unroll
        for (int i = 0; i < N; i++) out[i] = o.acc[i] * inv;
        store_vec<NV>(O + (int64_t)bh * hd, out, hd, lane);
    }
}

// ---- prefill, [batch, heads, seq, hd]; one warp per query row --------------
template <int NV>
__global__ void __launch_bounds__(kDecWarps * 32) prefill_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                                 const float* __restrict__ V, float* __restrict__ O, int seq, int hd,
                                                                 int causal) {
    constexpr int N = nvals<NV>();
    int bh = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int row = blockIdx.y * kDecWarps + warp;
    if (row >= seq) return;
    const float* Kb = K + (int64_t)bh * seq * hd;
    const float* Vb = V + (int64_t)bh * seq * hd;
    float q[N];
    load_vec<NV>(q, Q + ((int64_t)bh * seq + row) * hd, hd, lane);
    Online o;
    online_init(o);
    walk<NV>(o, q, Kb, Vb, hd, 0, causal ? row + 1 : seq, 1, hd, lane, rsqrtf((float)hd));
    float inv = o.l > 0.0f ? 1.0f / o.l : 0.0f;
    float out[N];
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = o.acc[i] * inv;
    store_vec<NV>(O + ((int64_t)bh * seq + row) * hd, out, hd, lane);
}

}  // namespace

#define ZB_DISPATCH_HD(hd, CALL)            \
    do {                                    \
        if ((hd) == 128) { CALL(1); }       \
        else if ((hd) == 256) { CALL(2); }  \
        else { CALL(0); }                   \
    } while (0)

ZB_API cudaError_t flash_decode_splitkv_f32(const float* Q, const float* K, const float* V, float* O, float* partial_O,
                                            float* partial_lse, int num_bh, int max_kv_len, int head_dim, int kv_len,
                                            const int* kv_len_ptr, int num_q_heads, int num_kv_heads, int chunk_size,
                                            cudaStream_t stream) {
    if (head_dim > kMaxHd || head_dim <= 0 || chunk_size <= 0 || num_kv_heads <= 0 || num_q_heads % num_kv_heads) return cudaErrorInvalidValue;
    if (num_bh <= 0) return cudaSuccess;
    int splits = (kv_len + chunk_size - 1) / chunk_size;
    if (splits < 1) splits = 1;
    if (splits > 65535) return cudaErrorInvalidValue;
    float* pM = partial_lse;
    float* pL = partial_lse + (int64_t)num_bh * splits;
    dim3 grid(num_bh, splits);
#define CALL(NV) decode_split_kernel<NV><<<grid, kDecWarps * 32, 0, stream>>>(Q, K, V, partial_O, pM, pL, max_kv_len, head_dim, kv_len, kv_len_ptr, num_q_heads, num_kv_heads, chunk_size, splits)
    ZB_DISPATCH_HD(head_dim, CALL);
#undef CALL
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    decode_reduce_kernel<<<num_bh, 128, 0, stream>>>(partial_O, pM, pL, O, head_dim, splits);
    return cudaGetLastError();
}

ZB_API cudaError_t flash_attention_decode_f32(const float* Q, const float* K, const float* V, float* O, int num_bh, int max_kv_len,
                                              int head_dim, int kv_len, const int* kv_len_ptr, int num_q_heads, int num_kv_heads,
                                              cudaStream_t stream) {
    if (head_dim > kMaxHd || head_dim <= 0 || num_kv_heads <= 0 || num_q_heads % num_kv_heads) return cudaErrorInvalidValue;
    if (num_bh <= 0) return cudaSuccess;
#define CALL(NV) decode_head_kernel<NV><<<num_bh, kDecWarps * 32, 0, stream>>>(Q, K, V, O, max_kv_len, head_dim, kv_len, kv_len_ptr, num_q_heads, num_kv_heads)
    ZB_DISPATCH_HD(head_dim, CALL);
#undef CALL
    return cudaGetLastError();
}

ZB_API cudaError_t flash_attention_forward_f32(const float* Q, const float* K, const float* V, float* O, int batch, int heads,
                                               int seq_len, int head_dim, int causal, cudaStream_t stream) {
    if (head_dim > kMaxHd || head_dim <= 0) return cudaErrorInvalidValue;
    if (batch * heads <= 0 || seq_len <= 0) return cudaSuccess;
    dim3 grid(batch * heads, zb::cdiv(seq_len, kDecWarps));
#define CALL(NV) prefill_kernel<NV><<<grid, kDecWarps * 32, 0, stream>>>(Q, K, V, O, seq_len, head_dim, causal)
    ZB_DISPATCH_HD(head_dim, CALL);
#undef CALL
    return cudaGetLastError();
}
