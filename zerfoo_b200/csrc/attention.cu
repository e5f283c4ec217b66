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
#include <cuda_fp16.h>
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

// ===========================================================================
// Decode attention stage (engine path): QK-norm + RoPE + KV append + split-KV
// flash decode + split merge in ONE launch.
//
// Replaces, per layer and token, rope_select + fused_qk_norm_rope / fused_rope +
// 2*nKV offset_memcpy + MatMulTransposeB + GPUFusedSoftmaxVMul (or
// flash_decode_splitkv + its reduce): grouped_query_attention.go:579-1121,
// generate/tensor_cache.go:205-262, flash_decode.cu:60-229.
//
// Layout: the cache is [nKV][max_seq][hd] per layer (per-KV-head contiguous, as
// TensorCache hands it to attention, tensor_cache.go:487-567), so the tile a CTA
// needs -- `chunk` consecutive positions of one KV head -- is one contiguous
// span and arrives with two TMA bulk copies.  One CTA per (KV head, split)
// serves all nQ/nKV query heads that share the KV head: K/V are read once, not
// once per query head (no Repeat, gqa.go:976-1047).  The token's own K/V row is
// rotated in the CTA that owns its position, used from shared memory and written
// to the cache by that CTA alone.  The last CTA of a KV head to finish (atomic
// ticket) merges the split partials.
// ===========================================================================
#include "zb_attn_tile.cuh"
#include "zb200.h"

namespace {

template <int EPL, int REP, int AW, typename KVT>
__global__ void __launch_bounds__(AW * 32) decode_attn_kernel(const AttnArgs p) {
    extern __shared__ __align__(128) uint8_t smraw[];
    __shared__ __align__(8) unsigned long long bar_storage;
    __shared__ int s_last;
    const uint32_t bar = smem_u32(&bar_storage);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();
    const int bz = blockIdx.z, split = blockIdx.y;
    const int pos = p.pos_ptr[bz];
    if (pos < 0 || pos >= p.max_seq) return;
    const int len = pos + 1;
    if (split * p.chunk >= len) return;
    if ((len + p.chunk - 1) / p.chunk > (int)gridDim.y) __trap();   // ZB_ATTN_SINGLE_TILE promise broken: fail loudly instead of dropping positions
    decode_attn_item<EPL, REP, AW, 0, KVT>(p, blockIdx.x, split, bz, pos, smraw, bar, 0u, &s_last);
}

template <int EPL, int REP, int AW, typename KVT = float>
cudaError_t launch_decode_attn_w(const AttnArgs& a, int batch, bool pdl, cudaStream_t stream) {
    // tile (K, V) is reused for the per-warp partial outputs: AW*REP <= 2*chunk because chunk >= 16, REP <= 8
    size_t floats = 2 * (size_t)a.chunk * a.hd + (size_t)REP * a.hd + (size_t)AW * a.hd + 2 * AW * REP;
    size_t smem = floats * 4;
    static DeviceOnce once;
    if (smem > 48 * 1024)
        if (cudaError_t e = once.ensure(smem, [&] { return cudaFuncSetAttribute(decode_attn_kernel<EPL, REP, AW, KVT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }))
            return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.nkv, a.max_splits, batch);
    cfg.blockDim = dim3(AW * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, decode_attn_kernel<EPL, REP, AW, KVT>, a);
}

// Warps per attention CTA: 4 by default; ZB_ATTN_WARPS=8|16 for long tiles (ZB_ATTN_CHUNK) -- tuning knob.
template <int EPL, int REP>
cudaError_t launch_decode_attn(const AttnArgs& a, int batch, bool pdl, cudaStream_t stream) {
    static const int aw_env = [] { const char* v = getenv("ZB_ATTN_WARPS"); return (v && v[0]) ? atoi(v) : 4; }();
    const int aw = a.warps > 0 ? a.warps : aw_env;
    if (a.kv_f16) {   // fp16 cache (generate/tensor_cache.go:224-238): half the KV bytes per step, f32 arithmetic
        if (aw >= 8 && 8 * REP <= 2 * a.chunk) return launch_decode_attn_w<EPL, REP, 8, __half>(a, batch, pdl, stream);
        return launch_decode_attn_w<EPL, REP, 4, __half>(a, batch, pdl, stream);
    }
    if (aw == 16 && 16 * REP <= 2 * a.chunk) return launch_decode_attn_w<EPL, REP, 16>(a, batch, pdl, stream);
    if (aw == 8 && 8 * REP <= 2 * a.chunk) return launch_decode_attn_w<EPL, REP, 8>(a, batch, pdl, stream);
    return launch_decode_attn_w<EPL, REP, 4>(a, batch, pdl, stream);
}

template <int EPL>
cudaError_t dispatch_rep(const AttnArgs& a, int rep, int batch, bool pdl, cudaStream_t s) {
    switch (rep) {
        case 1: return launch_decode_attn<EPL, 1>(a, batch, pdl, s);
        case 2: return launch_decode_attn<EPL, 2>(a, batch, pdl, s);
        case 3: return launch_decode_attn<EPL, 3>(a, batch, pdl, s);
        case 4: return launch_decode_attn<EPL, 4>(a, batch, pdl, s);
        case 8: return launch_decode_attn<EPL, 8>(a, batch, pdl, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

ZB_API int zb_decode_attn_f32(const zb_attn_args* a, int flags, zb_stream_t stream) {
    if (!a || a->head_dim <= 0 || a->n_kv <= 0 || a->n_q % a->n_kv || a->chunk < 16 || a->max_splits < 1 ||
        (a->max_splits * a->chunk < a->max_seq && !(flags & ZB_ATTN_SINGLE_TILE)))
        return cudaErrorInvalidValue;
    AttnArgs p{a->qkv, a->q_norm, a->k_norm, a->cos_tbl, a->sin_tbl, a->pos, a->k_cache, a->v_cache, a->out, a->part_o, a->part_ml, a->ticket,
               a->eps, (float)(1.0 / sqrt((double)a->head_dim)), a->head_dim, a->n_q, a->n_kv, a->max_seq, a->chunk, a->max_splits,
               a->block_table, a->max_blocks, a->page > 0 ? a->page : 16, a->qkv_stride, a->out_stride, a->warps,
               nullptr, 0, 0, 0, nullptr, a->window, a->window_on, a->kv_f16};
    const int batch = a->batch > 0 ? a->batch : 1;
    if (a->block_table && (a->chunk % p.page)) return cudaErrorInvalidValue;
    const int rep = a->n_q / a->n_kv;
    const bool pdl = (flags & 1) != 0;
    switch (a->head_dim) {
        case 32: return dispatch_rep<1>(p, rep, batch, pdl, (cudaStream_t)stream);
        case 64: return dispatch_rep<2>(p, rep, batch, pdl, (cudaStream_t)stream);
        case 128: return dispatch_rep<4>(p, rep, batch, pdl, (cudaStream_t)stream);
        case 256: return dispatch_rep<8>(p, rep, batch, pdl, (cudaStream_t)stream);
    }
    return cudaErrorInvalidValue;
}

// ===========================================================================
// Chunked prefill: a block of T prompt tokens of one sequence.
//   prefill_rope_append: per (token, head) QK-RMSNorm + half-split RoPE at position p0+i, K/V rows appended to the
//                        [n_kv][max_seq][hd] cache, rotated queries written to q_out [T][n_q*hd]
//   prefill_attn:        causal attention of every query row over cache rows [0, p0+i]; one warp per (head, query),
//                        K/V read straight from the cache (GQA: kv head = q head / rep)
// Replaces flash_attention_forward_f32's role for the prompt (flash.go:49-175, flash_attention.cu:43-153) on the cache layout
// of the decode path.  window > 0: the causal sliding-window mask of the reference's prompt pass (row i sees j <= i with
// i - j < window: grouped_query_attention.go:1074-1077,1395-1415); decode steps attend the whole cache.
// ===========================================================================
namespace {

__global__ void __launch_bounds__(128) prefill_rope_append_kernel(const float* __restrict__ qkv, int ld, const float* __restrict__ wq,
                                                                  const float* __restrict__ wk, const float* __restrict__ cos_tbl,
                                                                  const float* __restrict__ sin_tbl, int p0, float* __restrict__ q_out,
                                                                  float* __restrict__ kc, float* __restrict__ vc, float eps, int hd, int nq,
                                                                  int nkv, int max_seq) {
    extern __shared__ float xn[];
    __shared__ float red[32];
    const int i = blockIdx.x, head = blockIdx.y, pos = p0 + i;
    if (pos >= max_seq) return;
    const float* x = qkv + (size_t)i * ld + (size_t)head * hd;
    const int half = hd >> 1;
    if (head >= nq + nkv) {
        float* dst = vc + ((size_t)(head - nq - nkv) * max_seq + pos) * hd;
        for (int d = threadIdx.x; d < hd; d += blockDim.x) dst[d] = x[d];
        return;
    }
    const float* w = head < nq ? wq : wk;
    if (w) {
        float ss = 0.0f;
        for (int d = threadIdx.x; d < hd; d += blockDim.x) ss = fmaf(x[d], x[d], ss);
        ss = block_sum(ss, red);
        float s = (float)(1.0 / sqrt((double)(ss / (float)hd + eps)));
        for (int d = threadIdx.x; d < hd; d += blockDim.x) xn[d] = x[d] * s * w[d];
    } else {
        for (int d = threadIdx.x; d < hd; d += blockDim.x) xn[d] = x[d];
    }
    __syncthreads();
    float* o = head < nq ? q_out + (size_t)i * nq * hd + (size_t)head * hd : kc + ((size_t)(head - nq) * max_seq + pos) * hd;
    const float* cs = cos_tbl + (size_t)pos * half;
    const float* sn = sin_tbl + (size_t)pos * half;
    for (int d = threadIdx.x; d < half; d += blockDim.x) {
        float a = xn[d], b = xn[d + half], c = cs[d], s = sn[d];
        o[d] = a * c - b * s;
        o[d + half] = b * c + a * s;
    }
}

// Causal attention of a prompt chunk over the decode cache.  One warp owns a GQA group (REP query heads of one KV head) at QR
// consecutive prompt rows: every K / V row it loads feeds REP*QR dot products, which cuts the L2 traffic of the naive
// one-row-per-thread walk (flash_attention.cu:43-175) by that factor.  f32 throughout, online softmax per (head, row).
template <int NV, int REP, int QR>
__global__ void __launch_bounds__(kDecWarps * 32) prefill_attn_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                                      const float* __restrict__ V, float* __restrict__ O, int T, int p0,
                                                                      int hd, int nq, int nkv, int max_seq, float scale, int window) {
    constexpr int N = nvals<NV>(), G = REP * QR;
    const int kvh = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // long rows first: the grid's tail is then made of the cheap (short-prefix) rows
    const int i0 = ((int)gridDim.y - 1 - (int)blockIdx.y) * (kDecWarps * QR) + warp * QR;
    if (i0 >= T) return;
    const float* Kb = K + (size_t)kvh * max_seq * hd;
    const float* Vb = V + (size_t)kvh * max_seq * hd;
    float q[G][N], acc[G][N], m[G], l[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int row = min(i0 + g / REP, T - 1), h = kvh * REP + g % REP;
        load_vec<NV>(q[g], Q + ((size_t)row * nq + h) * hd, hd, lane);
        m[g] = -FLT_MAX; l[g] = 0.0f;
#pragma unroll
        for (int e = 0; e < N; e++) { q[g][e] *= scale; acc[g][e] = 0.0f; }
    }
    const int t_end = min(p0 + i0 + QR, p0 + T);
    const int t_begin = window > 0 ? max(0, p0 + i0 - window + 1) : 0;   // sliding window: row at position p sees (p - window, p]
    for (int t = t_begin; t < t_end; t++) {
        float kk[N], vv[N];
        load_vec<NV>(kk, Kb + (size_t)t * hd, hd, lane);
        load_vec<NV>(vv, Vb + (size_t)t * hd, hd, lane);
        float sc[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
            float d = 0.0f;
#pragma unroll
            for (int e = 0; e < N; e++) d = fmaf(q[g][e], kk[e], d);
            sc[g] = d;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int g = 0; g < G; g++) sc[g] += __shfl_xor_sync(0xffffffffu, sc[g], off);
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (t > p0 + i0 + g / REP) continue;  // causal mask (warp-uniform)
            if (window > 0 && p0 + i0 + g / REP - t >= window) continue;   // BuildCausalSlidingWindowMask: i - j < window
            const float mn = fmaxf(m[g], sc[g]);
            const float corr = __expf(m[g] - mn), p = __expf(sc[g] - mn);
            l[g] = l[g] * corr + p;
#pragma unroll
            for (int e = 0; e < N; e++) acc[g][e] = fmaf(p, vv[e], acc[g][e] * corr);
            m[g] = mn;
        }
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int row = i0 + g / REP, h = kvh * REP + g % REP;
        if (row >= T) continue;
        const float inv = l[g] > 0.0f ? 1.0f / l[g] : 0.0f;
        float out[N];
#pragma unroll
        for (int e = 0; e < N; e++) out[e] = acc[g][e] * inv;
        store_vec<NV>(O + ((size_t)row * nq + h) * hd, out, hd, lane);
    }
}

// ---- tensor-core variant: fp16 operands (q.k and p.v through mma.sync m16n8k16), f32 accumulate and softmax ---------
// A CTA = one query head x 64 prompt rows (4 warps x 16 rows).  K / V tiles of BK cache rows arrive as f32 through a two-stage
// cp.async ring (row strides padded to HD+8 / HD+4 floats: conflict-free fragment reads) and are rounded to fp16 while the
// B fragments are built; P stays in registers between the two MMAs (the accumulator layout of S is the A layout of P).
// Adjacent CTAs are the heads of one GQA group: they share the K / V tiles through L2.
__device__ __forceinline__ void pf_mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pf_pack(float a, float b) {
    __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float pf_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void pf_cp16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int HD, int BK>
__global__ void __launch_bounds__(128) prefill_attn_mma_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                               const float* __restrict__ V, float* __restrict__ O, int T, int p0, int nq,
                                                               int nkv, int max_seq, float qscale, int window) {
    constexpr int KS = HD + 8, VS = HD + 4, STAGE = BK * KS + BK * VS, NS = BK / 8, ND = HD / 8, NK = HD / 16;
    extern __shared__ __align__(16) float pf_sm[];
    const int h = blockIdx.x, kvh = h / (nq / nkv);
    const int row_cta = ((int)gridDim.y - 1 - (int)blockIdx.y) * 64;  // long prefixes first
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r0 = row_cta + warp * 16 + g, r1 = r0 + 8;
    const int rc0 = min(r0, T - 1), rc1 = min(r1, T - 1);
    uint32_t qf[NK][4];
    {
        const float* q0 = Q + ((size_t)rc0 * nq + h) * HD;
        const float* q1 = Q + ((size_t)rc1 * nq + h) * HD;
#pragma unroll
        for (int ks = 0; ks < NK; ks++) {
            const float2 a = *reinterpret_cast<const float2*>(q0 + 16 * ks + 2 * t), b = *reinterpret_cast<const float2*>(q1 + 16 * ks + 2 * t);
            const float2 c = *reinterpret_cast<const float2*>(q0 + 16 * ks + 8 + 2 * t), d = *reinterpret_cast<const float2*>(q1 + 16 * ks + 8 + 2 * t);
            qf[ks][0] = pf_pack(a.x * qscale, a.y * qscale); qf[ks][1] = pf_pack(b.x * qscale, b.y * qscale);
            qf[ks][2] = pf_pack(c.x * qscale, c.y * qscale); qf[ks][3] = pf_pack(d.x * qscale, d.y * qscale);
        }
    }
    float o[ND][4];
#pragma unroll
    for (int i = 0; i < ND; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.0f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    const int kv_cta = p0 + min(row_cta + 64, T);               // the CTA needs cache rows [0, kv_cta)
    const int kv_warp = p0 + min(row_cta + warp * 16 + 16, T);  // this warp needs [0, kv_warp)
    const int lim0 = p0 + rc0, lim1 = p0 + rc1;                 // row r attends to cache rows <= p0 + r
    // sliding window (prompt pass of the Mistral family, grouped_query_attention.go:1395-1415): ... and > p0 + r - window
    const int lo0 = window > 0 ? max(0, lim0 - window + 1) : 0, lo1 = window > 0 ? max(0, lim1 - window + 1) : 0;
    const int lo_warp = window > 0 ? max(0, p0 + row_cta + warp * 16 - window + 1) : 0;      // lowest bound among the warp's rows
    const int lo_warp_hi = window > 0 ? max(0, p0 + min(row_cta + warp * 16 + 15, T - 1) - window + 1) : 0;   // highest
    const int tl_first = window > 0 ? max(0, p0 + row_cta - window + 1) / BK : 0;             // tiles below every row's window are skipped
    const int n_tiles = (kv_cta + BK - 1) / BK;
    const float* Kb = K + (size_t)kvh * max_seq * HD;
    const float* Vb = V + (size_t)kvh * max_seq * HD;
    auto issue = [&](int tl) {
        float* ks_ = pf_sm + (tl & 1) * STAGE;
        float* vs_ = ks_ + BK * KS;
        for (int c = threadIdx.x; c < BK * (HD / 4); c += 128) {
            const int r = c / (HD / 4), col = (c % (HD / 4)) * 4, key = tl * BK + r;
            const bool ok = key < kv_cta;
            const size_t off = (size_t)(ok ? key : 0) * HD + col;
            pf_cp16(smem_u32(ks_ + r * KS + col), Kb + off, ok ? 16 : 0);   // rows past the prefix are zero-filled
            pf_cp16(smem_u32(vs_ + r * VS + col), Vb + off, ok ? 16 : 0);
        }
    };
    issue(tl_first);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int tl = tl_first; tl < n_tiles; tl++) {
        if (tl + 1 < n_tiles) issue(tl + 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        if (tl * BK < kv_warp && tl * BK + BK > lo_warp) {
            const float* ks_ = pf_sm + (tl & 1) * STAGE;
            const float* vs_ = ks_ + BK * KS;
            float sc[NS][4];
#pragma unroll
            for (int n = 0; n < NS; n++) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.0f;
#pragma unroll
            for (int ks = 0; ks < NK; ks++) {
#pragma unroll
                for (int n = 0; n < NS; n++) {
                    const float* kr = ks_ + (n * 8 + g) * KS + 16 * ks + 2 * t;
                    const float2 x = *reinterpret_cast<const float2*>(kr), y = *reinterpret_cast<const float2*>(kr + 8);
                    pf_mma16816(sc[n], qf[ks], pf_pack(x.x, x.y), pf_pack(y.x, y.y));
                }
            }
            if (tl * BK + BK - 1 > p0 + row_cta + warp * 16) {  // the diagonal tile(s): causal mask
#pragma unroll
                for (int n = 0; n < NS; n++) {
                    const int key = tl * BK + n * 8 + 2 * t;
                    if (key > lim0) sc[n][0] = -INFINITY;
                    if (key + 1 > lim0) sc[n][1] = -INFINITY;
                    if (key > lim1) sc[n][2] = -INFINITY;
                    if (key + 1 > lim1) sc[n][3] = -INFINITY;
                }
            }
            if (tl * BK < lo_warp_hi) {   // the tile(s) the lower edge of the window cuts through
#pragma unroll
                for (int n = 0; n < NS; n++) {
                    const int key = tl * BK + n * 8 + 2 * t;
                    if (key < lo0) sc[n][0] = -INFINITY;
                    if (key + 1 < lo0) sc[n][1] = -INFINITY;
                    if (key < lo1) sc[n][2] = -INFINITY;
                    if (key + 1 < lo1) sc[n][3] = -INFINITY;
                }
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int n = 0; n < NS; n++) {
                mx0 = fmaxf(mx0, fmaxf(sc[n][0], sc[n][1]));
                mx1 = fmaxf(mx1, fmaxf(sc[n][2], sc[n][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            // without a window the maxima are finite from tile 0 on (cache row 0 is visible to every row); with one a row may
            // have seen nothing yet: exponentiate against 0 then, every term is exp2(-inf) = 0
            const float mx0n = fmaxf(m0, mx0), mx1n = fmaxf(m1, mx1);
            const float mn0 = mx0n == -INFINITY ? 0.0f : mx0n, mn1 = mx1n == -INFINITY ? 0.0f : mx1n;
            const float c0 = pf_ex2(m0 - mn0), c1 = pf_ex2(m1 - mn1);
            m0 = mx0n; m1 = mx1n;
            float s0 = 0.0f, s1 = 0.0f;
            uint32_t pa[NS / 2][4];
#pragma unroll
            for (int n = 0; n < NS; n++) {
                const float e0 = pf_ex2(sc[n][0] - mn0), e1 = pf_ex2(sc[n][1] - mn0), e2 = pf_ex2(sc[n][2] - mn1), e3 = pf_ex2(sc[n][3] - mn1);
                s0 += e0 + e1; s1 += e2 + e3;
                pa[n >> 1][(n & 1) * 2] = pf_pack(e0, e1);
                pa[n >> 1][(n & 1) * 2 + 1] = pf_pack(e2, e3);
            }
            l0 = l0 * c0 + s0; l1 = l1 * c1 + s1;
#pragma unroll
            for (int d = 0; d < ND; d++) { o[d][0] *= c0; o[d][1] *= c0; o[d][2] *= c1; o[d][3] *= c1; }
#pragma unroll
            for (int kk = 0; kk < NS / 2; kk++) {
#pragma unroll
                for (int d = 0; d < ND; d++) {
                    const float* vr = vs_ + (16 * kk + 2 * t) * VS + d * 8 + g;
                    pf_mma16816(o[d], pa[kk], pf_pack(vr[0], vr[VS]), pf_pack(vr[8 * VS], vr[9 * VS]));
                }
            }
        }
        __syncthreads();
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.0f ? 1.0f / l0 : 0.0f, i1 = l1 > 0.0f ? 1.0f / l1 : 0.0f;
#pragma unroll
    for (int d = 0; d < ND; d++) {
        if (r0 < T) *reinterpret_cast<float2*>(O + ((size_t)r0 * nq + h) * HD + d * 8 + 2 * t) = make_float2(o[d][0] * i0, o[d][1] * i0);
        if (r1 < T) *reinterpret_cast<float2*>(O + ((size_t)r1 * nq + h) * HD + d * 8 + 2 * t) = make_float2(o[d][2] * i1, o[d][3] * i1);
    }
}

template <int HD>
cudaError_t launch_prefill_attn_mma(const float* q_rot, const float* kc, const float* vc, float* out, int tokens, int p0, int nq, int nkv,
                                    int max_seq, float scale, int window, cudaStream_t s) {
    constexpr int BK = 32;
    constexpr size_t smem = 2 * (size_t)(BK * (HD + 8) + BK * (HD + 4)) * sizeof(float);
    static DeviceOnce once;
    if (smem > 48 * 1024)
        if (cudaError_t e = once.ensure(smem, [&] { return cudaFuncSetAttribute(prefill_attn_mma_kernel<HD, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }))
            return e;
    prefill_attn_mma_kernel<HD, BK><<<dim3(nq, (tokens + 63) / 64), 128, smem, s>>>(q_rot, kc, vc, out, tokens, p0, nq, nkv, max_seq,
                                                                                   scale * 1.4426950408889634f, window);
    return cudaGetLastError();
}

template <int NV>
cudaError_t launch_prefill_attn(const float* q_rot, const float* kc, const float* vc, float* out, int tokens, int p0, int hd, int nq, int nkv,
                                int max_seq, float scale, int window, cudaStream_t s) {
    const int rep = nq / nkv;
#define ZB_PF(REP, QR)                                                                                                              \
    prefill_attn_kernel<NV, REP, QR><<<dim3(nkv, (tokens + kDecWarps * QR - 1) / (kDecWarps * QR)), kDecWarps * 32, 0, s>>>(      \
        q_rot, kc, vc, out, tokens, p0, hd, nq, nkv, max_seq, scale, window)
    switch (rep) {
        case 1: ZB_PF(1, 4); break;
        case 2: ZB_PF(2, 2); break;
        case 3: ZB_PF(3, 2); break;
        case 4: ZB_PF(4, 2); break;
        case 8: ZB_PF(8, 1); break;
        default: return cudaErrorInvalidValue;
    }
#undef ZB_PF
    return cudaGetLastError();
}

}  // namespace

ZB_API int zb_prefill_attn_f32(const float* qkv, int ld_qkv, const float* q_norm, const float* k_norm, const float* cos_tbl,
                               const float* sin_tbl, int p0, int tokens, float* q_rot, float* k_cache, float* v_cache, float* out, float eps,
                               int head_dim, int n_q, int n_kv, int max_seq, int window, int flags, zb_stream_t stream) {
    if (tokens <= 0 || head_dim <= 0 || head_dim > kMaxHd || (head_dim & 1) || n_kv <= 0 || n_q % n_kv || p0 < 0 || p0 + tokens > max_seq || window < 0)
        return cudaErrorInvalidValue;
    cudaStream_t s = (cudaStream_t)stream;
    prefill_rope_append_kernel<<<dim3(tokens, n_q + 2 * n_kv), 128, head_dim * sizeof(float), s>>>(
        qkv, ld_qkv, q_norm, k_norm, cos_tbl, sin_tbl, p0, q_rot, k_cache, v_cache, eps, head_dim, n_q, n_kv, max_seq);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const float scale = (float)(1.0 / sqrt((double)head_dim));
    if (!(flags & ZB_PREFILL_ATTN_F32)) {
        if (head_dim == 128) return launch_prefill_attn_mma<128>(q_rot, k_cache, v_cache, out, tokens, p0, n_q, n_kv, max_seq, scale, window, s);
        if (head_dim == 64) return launch_prefill_attn_mma<64>(q_rot, k_cache, v_cache, out, tokens, p0, n_q, n_kv, max_seq, scale, window, s);
        if (head_dim == 32) return launch_prefill_attn_mma<32>(q_rot, k_cache, v_cache, out, tokens, p0, n_q, n_kv, max_seq, scale, window, s);
    }
#define CALL(NV) return launch_prefill_attn<NV>(q_rot, k_cache, v_cache, out, tokens, p0, head_dim, n_q, n_kv, max_seq, scale, window, s)
    ZB_DISPATCH_HD(head_dim, CALL);
#undef CALL
    return cudaErrorInvalidValue;
}
