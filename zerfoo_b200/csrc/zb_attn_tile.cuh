// One (KV head, split) work item of the fused decode-attention stage, shared by the per-layer kernel (attention.cu:
// decode_attn_kernel) and the persistent whole-token kernel (decode_mega.cu): QK-RMSNorm + half-split RoPE of the token's
// query heads, KV append by the item that owns the token's position, online softmax over the item's tile of the
// [n_kv][max_seq][hd] cache (two TMA bulk copies), merge of the warps, and -- when the context spans several splits -- the
// ticketed merge of the split partials by the last item of the KV head to finish.
//
// Reference semantics replaced: grouped_query_attention.go:579-1121 (QK-norm, RoPE, Repeat + MatMulTransposeB +
// GPUFusedSoftmaxVMul), generate/tensor_cache.go:205-262 (append at the device counter), flash_decode.cu:60-229.
//
// Every value another CTA produced earlier in the same launch (the qkv projections, the split partials) is read with
// ld.global.cg: inside a persistent kernel the L1 is not invalidated between the layers.
#pragma once
#include <float.h>

#include "zb_stream.cuh"

namespace zb {

struct AttnArgs {
    const float* qkv;
    const float* wq;
    const float* wk;
    const float* cos_tbl;
    const float* sin_tbl;
    const int* pos_ptr;
    void* kc;                 // [n_kv][max_seq][hd] (or the paged pool) of f32, or of fp16 when kv_f16 (generate/tensor_cache.go:224-238)
    void* vc;
    float* out;
    float* part_o;
    float* part_ml;
    int* ticket;
    float eps, scale;
    int hd, nq, nkv, max_seq, chunk, max_splits;
    // batched / paged extension (grid.z = sequence): per-sequence strides and an optional block table
    const int* block_table;   // [batch][max_blocks] physical page ids, or nullptr for the contiguous [n_kv][max_seq][hd] cache
    int max_blocks, page;     // page = positions per block (16, generate/generator.go:238); pool layout [block][n_kv][page][hd]
    int qkv_stride, out_stride;
    int warps;
    // persistent kernel only (decode_mega.cu): qkv and out as flagged vectors (zb_mega.cuh MegaVec); epochs are call arguments
    const uint2* qkv_ll;
    int qkv_planes, qkv_ll_stride, qkv_tag_op;
    uint2* out_ll;
    // sliding window of the prompt pass (zb_attn_args.window / window_on)
    int window;
    const int* window_on;
    int kv_f16;
};

// ---- flagged vectors: (value bits, epoch) pairs, polled until the epoch matches ---------------------------------------------
__device__ __forceinline__ void st_pair(uint2* p, float v, uint32_t epoch) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
__device__ __forceinline__ void st_pair2(uint2* p, float v0, float v1, uint32_t epoch) {   // p 16-byte aligned
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(v0)), "r"(epoch), "r"(__float_as_uint(v1)),
                 "r"(epoch)
                 : "memory");
}
// elements lane, lane + 32, ... of a row of hd = 32 EPL flagged values, planes summed in plane order
template <int EPL>
__device__ __forceinline__ void ll_row(const uint2* row, int planes, int stride, int lane, uint32_t epoch, float (&out)[EPL]) {
    for (int pl = 0; pl < planes; pl++) {
        const uint2* r = row + (size_t)pl * stride + lane;
        uint32_t v[EPL], f[EPL], spins = 0;
        bool ok;
        do {
#pragma unroll
            for (int e = 0; e < EPL; e++) asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[e]), "=r"(f[e]) : "l"(r + 32 * e) : "memory");
            ok = true;
#pragma unroll
            for (int e = 0; e < EPL; e++) ok = ok && f[e] == epoch;
            if (++spins > (1u << 24)) __trap();   // a lost producer traps instead of hanging the GPU
        } while (!ok);
#pragma unroll
        for (int e = 0; e < EPL; e++) out[e] = pl == 0 ? __uint_as_float(v[e]) : out[e] + __uint_as_float(v[e]);
    }
}


// One warp: per-head RMSNorm (optional) + half-split RoPE of `src` (global, L2) into `dst` (shared), using `tmp` (shared, hd floats).
// epoch != 0: the row comes from the flagged vector `ll` (element index `ll_elem`) instead of `src`.
template <int EPL>
__device__ __forceinline__ void norm_rope_warp(const float* __restrict__ src, const float* __restrict__ w, const float* __restrict__ cs,
                                               const float* __restrict__ sn, float* tmp, float* dst, int hd, float eps, int lane,
                                               const uint2* ll = nullptr, int ll_planes = 0, int ll_stride = 0, uint32_t epoch = 0u) {
    int half = hd >> 1;
    if (epoch) {
        float t[EPL];
        ll_row<EPL>(ll, ll_planes, ll_stride, lane, epoch, t);
#pragma unroll
        for (int e = 0; e < EPL; e++) tmp[lane + 32 * e] = t[e];
    } else {
        for (int d = lane; d < hd; d += 32) tmp[d] = __ldcg(src + d);
    }
    if (w) {
        float ss = 0.0f;
        for (int d = lane; d < hd; d += 32) {
            const float v = tmp[d];
            ss = fmaf(v, v, ss);
        }
        ss = warp_sum(ss);
        float s = (float)(1.0 / sqrt((double)(ss / (float)hd + eps)));
        for (int d = lane; d < hd; d += 32) tmp[d] = tmp[d] * s * w[d];
    }
    __syncwarp();
    for (int d = lane; d < half; d += 32) {
        float a = tmp[d], b = tmp[d + half], c = cs[d], s = sn[d];
        dst[d] = a * c - b * s;
        dst[d + half] = b * c + a * s;
    }
    __syncwarp();
}

// lane owns EPL head-dim elements: EPL <= 4 -> contiguous [lane*EPL, +EPL); EPL == 8 -> two float4 at lane*4 and 128 + lane*4
template <int EPL>
__device__ __forceinline__ void ld_row(float (&v)[EPL], const float* row, int lane) {
    if (EPL == 8) {
        float4 a = *reinterpret_cast<const float4*>(row + lane * 4), b = *reinterpret_cast<const float4*>(row + 128 + lane * 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if (EPL == 4) {
        float4 a = *reinterpret_cast<const float4*>(row + lane * 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else if (EPL == 2) {
        float2 a = *reinterpret_cast<const float2*>(row + lane * 2);
        v[0] = a.x; v[1] = a.y;
    } else {
        v[0] = row[lane];
    }
}
// the same ownership out of an fp16 cache tile (KV stored as fp16, arithmetic in f32)
template <int EPL>
__device__ __forceinline__ void ld_row(float (&v)[EPL], const __half* row, int lane) {
    if (EPL == 8) {
        const uint2 a = *reinterpret_cast<const uint2*>(row + lane * 4), b = *reinterpret_cast<const uint2*>(row + 128 + lane * 4);
        const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
        const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&b.x)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&b.y));
        v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y; v[4] = b0.x; v[5] = b0.y; v[6] = b1.x; v[7] = b1.y;
    } else if (EPL == 4) {
        const uint2 a = *reinterpret_cast<const uint2*>(row + lane * 4);
        const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
        v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y;
    } else if (EPL == 2) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(row + lane * 2));
        v[0] = a.x; v[1] = a.y;
    } else {
        v[0] = __half2float(row[lane]);
    }
}
__device__ __forceinline__ float kv_round(float v, const float*) { return v; }
__device__ __forceinline__ float kv_round(float v, const __half*) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ void kv_store(float* p, float v) { *p = v; }
__device__ __forceinline__ void kv_store(__half* p, float v) { *p = __float2half_rn(v); }

template <int EPL>
__device__ __forceinline__ void st_row(float* row, const float (&v)[EPL], int lane) {
    if (EPL == 8) {
        *reinterpret_cast<float4*>(row + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(row + 128 + lane * 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if (EPL == 4) {
        *reinterpret_cast<float4*>(row + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (EPL == 2) {
        *reinterpret_cast<float2*>(row + lane * 2) = make_float2(v[0], v[1]);
    } else {
        row[lane] = v[0];
    }
}

template <int EPL>
__device__ __forceinline__ void st_row_ll(uint2* row, const float (&v)[EPL], int lane, uint32_t epoch) {
    if (EPL == 8) {
        st_pair2(row + lane * 4, v[0], v[1], epoch); st_pair2(row + lane * 4 + 2, v[2], v[3], epoch);
        st_pair2(row + 128 + lane * 4, v[4], v[5], epoch); st_pair2(row + 128 + lane * 4 + 2, v[6], v[7], epoch);
    } else if (EPL == 4) {
        st_pair2(row + lane * 4, v[0], v[1], epoch); st_pair2(row + lane * 4 + 2, v[2], v[3], epoch);
    } else if (EPL == 2) {
        st_pair2(row + lane * 2, v[0], v[1], epoch);
    } else {
        st_pair(row + lane, v[0], epoch);
    }
}

// Barrier among the AW warps that work on the item: the whole CTA (__syncthreads) in the per-layer kernel, a named
// barrier when the item runs on a warp group of a larger persistent CTA.
template <int AW, int BAR_ID>
__device__ __forceinline__ void attn_group_sync() {
    if (BAR_ID == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR_ID), "n"(AW * 32) : "memory");
}

// floats of shared memory one item needs (K tile, V tile, rotated queries, per-warp scratch, per-warp m / l)
__host__ __device__ inline size_t attn_item_floats(int chunk, int hd, int rep, int aw) {
    return 2 * (size_t)chunk * hd + (size_t)rep * hd + (size_t)aw * hd + 2 * (size_t)aw * rep;
}

// Threads 0 .. AW*32-1 of the group call this uniformly.  `pos` is the token's position (kv_len = pos + 1), `split` < nsplits.
// `bar` is an initialised mbarrier (count 1) used once per call with phase parity `parity`.
template <int EPL, int REP, int AW, int BAR_ID, typename KVT = float>
__device__ __forceinline__ void decode_attn_item(const AttnArgs& p, int kvh, int split, int bz, int pos, uint8_t* smraw, uint32_t bar,
                                                 uint32_t parity, int* s_last, uint32_t epoch_in = 0u, uint32_t epoch_out = 0u) {
    const int hd = p.hd;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the two cache tiles sit in the space of two f32 tiles whatever KVT is: the scratch behind them keeps its place
    KVT* sK = reinterpret_cast<KVT*>(smraw);                 // [chunk][hd]
    KVT* sV = sK + (size_t)p.chunk * hd;                     // [chunk][hd]
    float* sQ = reinterpret_cast<float*>(smraw) + 2 * (size_t)p.chunk * hd;   // [REP][hd]
    KVT* const kcache = static_cast<KVT*>(p.kc);
    KVT* const vcache = static_cast<KVT*>(p.vc);
    float* sT = sQ + (size_t)REP * hd;                       // [AW][hd] scratch
    float* sM = sT + (size_t)AW * hd;                        // [AW][REP] m, then l
    const float* qkv = p.qkv + (size_t)bz * p.qkv_stride;
    float* outp = p.out + (size_t)bz * p.out_stride;
    float* part_o = p.part_o + (size_t)bz * p.nq * p.max_splits * hd;
    float* part_ml = p.part_ml + (size_t)bz * 2 * p.nq * p.max_splits;
    int* ticket = p.ticket + (size_t)bz * p.nkv;
    const int* btab = p.block_table ? p.block_table + (size_t)bz * p.max_blocks : nullptr;
    const int len = pos + 1, t0 = split * p.chunk;
    const int t1 = min(t0 + p.chunk, len), n = t1 - t0;
    const int nsplits = (len + p.chunk - 1) / p.chunk;
    // prompt pass of a sliding-window model: cache rows below `lo` are masked (their weight is an exact 0 in the reference)
    const int lo = (p.window > 0 && p.window_on && *p.window_on && len > p.window) ? len - p.window : 0;
    const size_t head_base = (size_t)kvh * p.max_seq * hd;
    // address of cache row `t` of this KV head: contiguous cache, or page table lookup (PagedKVCache, generate/paged_kv.go:74-136)
    auto row_off = [&](int t) -> size_t {
        if (!btab) return head_base + (size_t)t * hd;
        return (((size_t)btab[t / p.page] * p.nkv + kvh) * p.page + (size_t)(t % p.page)) * hd;
    };
    if (threadIdx.x == 0) {
        uint32_t bytes = (uint32_t)n * hd * (uint32_t)sizeof(KVT);
        fence_proxy_async();   // a previous item's generic-proxy reads of the tile precede these async-proxy writes
        mbar_expect_tx(bar, 2 * bytes);
        if (!btab) {
            bulk_g2s(smem_u32(sK), kcache + row_off(t0), bytes, bar);
            bulk_g2s(smem_u32(sV), vcache + row_off(t0), bytes, bar);
        } else {  // chunk is a multiple of the page size: one bulk copy per page and tensor
            for (int t = t0; t < t1; t += p.page) {
                uint32_t pb = (uint32_t)min(p.page, t1 - t) * hd * (uint32_t)sizeof(KVT);
                bulk_g2s(smem_u32(sK + (size_t)(t - t0) * hd), kcache + row_off(t), pb, bar);
                bulk_g2s(smem_u32(sV + (size_t)(t - t0) * hd), vcache + row_off(t), pb, bar);
            }
        }
    }
    const int half = hd >> 1;
    const float* cs = p.cos_tbl + (size_t)pos * half;
    const float* sn = p.sin_tbl + (size_t)pos * half;
    for (int r = warp; r < REP; r += AW)
        norm_rope_warp<EPL>(qkv + (size_t)(kvh * REP + r) * hd, p.wq, cs, sn, sT + warp * hd, sQ + r * hd, hd, p.eps, lane,
                            p.qkv_ll + (size_t)(kvh * REP + r) * hd, p.qkv_planes, p.qkv_ll_stride, epoch_in);
    mbar_wait(bar, parity);
    attn_group_sync<AW, BAR_ID>();
    if (pos >= t0 && pos < t1) {  // this item owns the token's position: rotate K, take V, publish both (rounded to the cache's type)
        KVT* krow = sK + (size_t)(pos - t0) * hd;
        KVT* vrow = sV + (size_t)(pos - t0) * hd;
        if (warp == 0) {
            float* kf = sT + hd;   // warp 1's scratch row: idle while warp 1 copies V (AW >= 2)
            norm_rope_warp<EPL>(qkv + (size_t)(p.nq + kvh) * hd, p.wk, cs, sn, sT, kf, hd, p.eps, lane,
                                p.qkv_ll + (size_t)(p.nq + kvh) * hd, p.qkv_planes, p.qkv_ll_stride, epoch_in);
            const size_t ro = row_off(pos);
            for (int d = lane; d < hd; d += 32) {
                kv_store(krow + d, kf[d]);
                kv_store(kcache + ro + d, kf[d]);
            }
        } else if (warp == 1) {
            const float* v = qkv + (size_t)(p.nq + p.nkv + kvh) * hd;
            const size_t ro = row_off(pos);
            if (epoch_in) {
                float t[EPL];
                ll_row<EPL>(p.qkv_ll + (size_t)(p.nq + p.nkv + kvh) * hd, p.qkv_planes, p.qkv_ll_stride, lane, epoch_in, t);
#pragma unroll
                for (int e = 0; e < EPL; e++) {
                    kv_store(vrow + lane + 32 * e, t[e]);
                    kv_store(vcache + ro + lane + 32 * e, t[e]);
                }
            } else {
                for (int d = lane; d < hd; d += 32) {
                    float t = __ldcg(v + d);
                    kv_store(vrow + d, t);
                    kv_store(vcache + ro + d, t);
                }
            }
        }
        attn_group_sync<AW, BAR_ID>();
    }
    // ---- online softmax over this split: warp w takes positions w, w+AW, ...
    float q[REP][EPL], acc[REP][EPL], m[REP], l[REP];
#pragma unroll
    for (int r = 0; r < REP; r++) {
        ld_row<EPL>(q[r], sQ + r * hd, lane);
        m[r] = -FLT_MAX;
        l[r] = 0.0f;
#pragma unroll
        for (int e = 0; e < EPL; e++) acc[r][e] = 0.0f;
    }
    for (int t = warp; t < n; t += AW) {
        if (t0 + t < lo) continue;
        float kv[EPL], vv[EPL];
        ld_row<EPL>(kv, sK + (size_t)t * hd, lane);
        ld_row<EPL>(vv, sV + (size_t)t * hd, lane);
#pragma unroll
        for (int r = 0; r < REP; r++) {
            float s = 0.0f;
#pragma unroll
            for (int e = 0; e < EPL; e++) s = fmaf(q[r][e], kv[e], s);
            s = warp_sum(s) * p.scale;
            float mn = fmaxf(m[r], s);
            float corr = __expf(m[r] - mn), pe = __expf(s - mn);
            l[r] = l[r] * corr + pe;
#pragma unroll
            for (int e = 0; e < EPL; e++) acc[r][e] = fmaf(pe, vv[e], acc[r][e] * corr);
            m[r] = mn;
        }
    }
    // ---- merge the warps of the group (through shared memory; sK is dead now)
    attn_group_sync<AW, BAR_ID>();
    float* sAcc = reinterpret_cast<float*>(smraw);  // [AW][REP][hd] over the (dead) cache tiles
    float* sL = sM + AW * REP;
#pragma unroll
    for (int r = 0; r < REP; r++) {
        st_row<EPL>(sAcc + (size_t)(warp * REP + r) * hd, acc[r], lane);
        if (lane == 0) { sM[warp * REP + r] = m[r]; sL[warp * REP + r] = l[r]; }
    }
    attn_group_sync<AW, BAR_ID>();
    for (int r = warp; r < REP; r += AW) {
        float mm = -FLT_MAX;
        for (int w = 0; w < AW; w++) mm = fmaxf(mm, sM[w * REP + r]);
        float ll = 0.0f, o[EPL];
#pragma unroll
        for (int e = 0; e < EPL; e++) o[e] = 0.0f;
        for (int w = 0; w < AW; w++) {
            float lw = sL[w * REP + r];
            float c = lw > 0.0f ? __expf(sM[w * REP + r] - mm) : 0.0f;
            ll += lw * c;
            float v[EPL];
            ld_row<EPL>(v, sAcc + (size_t)(w * REP + r) * hd, lane);
#pragma unroll
            for (int e = 0; e < EPL; e++) o[e] = fmaf(v[e], c, o[e]);
        }
        const int h = kvh * REP + r;
        if (nsplits == 1) {
            float inv = ll > 0.0f ? 1.0f / ll : 0.0f;
#pragma unroll
            for (int e = 0; e < EPL; e++) o[e] *= inv;
            if (epoch_out) st_row_ll<EPL>(p.out_ll + (size_t)h * hd, o, lane, epoch_out);
            else st_row<EPL>(outp + (size_t)h * hd, o, lane);
        } else {
            size_t slot = (size_t)h * p.max_splits + split;
            st_row<EPL>(part_o + slot * hd, o, lane);
            if (lane == 0) { part_ml[2 * slot] = mm; part_ml[2 * slot + 1] = ll; }
        }
    }
    if (nsplits == 1) {
        attn_group_sync<AW, BAR_ID>();   // the tile is reused by the caller's next item
        return;
    }
    // ---- the last item of this KV head merges the splits (threadfence reduction)
    __threadfence();
    attn_group_sync<AW, BAR_ID>();
    if (threadIdx.x == 0) {
        int old = atomicAdd(ticket + kvh, 1);
        *s_last = (old == nsplits - 1);
        if (*s_last) ticket[kvh] = 0;  // re-arm for the next launch
    }
    attn_group_sync<AW, BAR_ID>();
    const bool last = *s_last != 0;
    attn_group_sync<AW, BAR_ID>();       // s_last is rewritten by the caller's next item
    if (!last) return;
    __threadfence();
    for (int r = warp; r < REP; r += AW) {
        const int h = kvh * REP + r;
        const size_t base = (size_t)h * p.max_splits;
        float mm = -FLT_MAX;
        for (int s = 0; s < nsplits; s++)
            if (__ldcg(part_ml + 2 * (base + s) + 1) > 0.0f) mm = fmaxf(mm, __ldcg(part_ml + 2 * (base + s)));
        float ll = 0.0f, o[EPL];
#pragma unroll
        for (int e = 0; e < EPL; e++) o[e] = 0.0f;
        for (int s = 0; s < nsplits; s++) {
            float ls = __ldcg(part_ml + 2 * (base + s) + 1);
            if (ls > 0.0f) {
                float c = __expf(__ldcg(part_ml + 2 * (base + s)) - mm);
                ll += ls * c;
                const float* po = part_o + (base + s) * hd;
#pragma unroll
                for (int e = 0; e < EPL; e++) {
                    int d = EPL == 8 ? (e < 4 ? lane * 4 + e : 128 + lane * 4 + (e - 4)) : lane * EPL + e;
                    o[e] = fmaf(__ldcg(po + d), c, o[e]);
                }
            }
        }
        float inv = ll > 0.0f ? 1.0f / ll : 0.0f;
#pragma unroll
        for (int e = 0; e < EPL; e++) o[e] *= inv;
        if (epoch_out) st_row_ll<EPL>(p.out_ll + (size_t)h * hd, o, lane, epoch_out);
        else st_row<EPL>(outp + (size_t)h * hd, o, lane);
    }
}

}  // namespace zb
