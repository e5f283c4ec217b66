// Tensor-core fused dequant-GEMV for batch-1 decode (sm_100a): y[M] = deq(W[M,K]) . x[K], x = prologue(...).
//
// Why tensor cores at batch 1.  At B200 rates (6.5 TB/s over 148 SMs) a 4-bit GEMV has an issue budget of ~3.2 lane
// instructions per weight; the CUDA-core dequant of gemv_stream.cu needs ~3.8 (PRMT + FADD2 + FFMA2 per weight) and tops
// out near half the HBM roofline.  Here the int->float conversion AND the multiply-accumulate run on the tensor pipe.
// Default (integer) path, mma.sync.m16n8k32.s32.u8.s8:
//   * a masked quant word (w & 0x0F0F0F0F) IS four u8 operands -- one LOP3 per four weights;
//   * per 256 weights x is scaled by a power of two (warp-local max) to a 32-bit fixed-point number and split into four
//     balanced base-256 digits, one per B column: products and sums are exact integers, block scales are applied with
//     IMAD, the Q4_K min term sum_s m_s * sum(x_s) is one more MMA, the Q6_K / Q4_0 offsets enter as the C operand;
//     only the final per-super-block combination is rounded (f32);
// f16 path (ZB_MMA_I8=0, mma.sync.m16n8k16): nibble pairs as exact fp16 subnormals n * 2^-24 | n * 2^-20, x as three fp16
//   terms; kept for comparison -- slower, and one HMMA aligns its addends to the largest and truncates ~17 bits below.
// Work unit: a "block-tile" = 16 rows x one 256-weight super-block (Q4_K 2304 B, Q5_K 2816 B, Q6_K 3360 B) or x four Q4_0
// blocks (1152 B), laid out at upload time so that every lane's operand bytes are conflict-free LDS.128 groups (pure byte
// permutation of the GGUF blocks: dequantised values stay bit-exact).  Block-tiles are numbered (row_tile * units_per_row +
// unit) = their order in memory; CTA c owns the contiguous range [c*q, (c+1)*q), warp w of the CTA a contiguous run of it,
// streamed 1-4 tiles at a time through a private TMA ring (cp.async.bulk + mbarrier, first fill issued before
// griddepcontrol.wait so it overlaps the previous kernel; MoE launches choose the expert on the device and fill after it).  Row tiles
// that straddle CTAs are finished by the CTA that owns their first part: the others push (value, flag) pairs as single
// 8-byte stores (the data is its own flag: no fence, no atomic), the owner polls them in part order (deterministic).
// One CTA of 16 warps per SM; the fused prologue runs in registers (warp w builds 256-element blocks w, w+16, ... of x).
//
// Reference semantics replaced: Engine.MatMul on Q4_K / Q5_K / Q6_K / Q4_0 storage (gemv_q4k.cu:68-160, dequant spec
// :14-21,38-56; gemv_q5k.cu:15-23,68-177; gemv_q6k.cu:11-25,45-152; gemm_q4.cu:1-12,89-96) plus the fused providers around it
// (fused_add_rmsnorm.cu:17-84, fused_norm_add.cu:11-81, fused_swiglu.cu:11-34) and the expert indirection of
// layers/core/moe.go:110-146.
#include <stdlib.h>
#include <string.h>

#include "zb_prologue.cuh"
#include "zb200.h"

#include "zb_mma_tiles.cuh"

namespace {


// MIX: the activation is a sum of mix_n weighted vectors (MoE combine, K-slab partial sums).  A separate instantiation: with
// the combine loop compiled in, the common kernel needs 122 registers instead of 98 and every in-step launch gets ~0.6 us slower.
template <int TYPE, int I8, bool MIX = false>
__global__ void __launch_bounds__(kMT, 1) gemv_mma_kernel(const uint8_t* __restrict__ wm, int M, int K, const MGeom g, const Prologue p,
                                                          float* __restrict__ y, int pairs, uint2* __restrict__ gpart,
                                                          unsigned long long* __restrict__ trace, const MSel ms) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float red[32];
    constexpr int BT = bt_bytes(TYPE);
    uint4* xf = reinterpret_cast<uint4*>(smem + g.xf_off);
    uint32_t* xm = reinterpret_cast<uint32_t*>(smem + g.xm_off);   // Q6_K: float offsets; lanes with t != 0 read the zero block behind it
    float* xinv = reinterpret_cast<float*>(smem + g.xinv_off);
    float* part = reinterpret_cast<float*>(smem + g.part_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // optional phase timeline (ZB_MMA_TRACE, tools/gemv_trace.py): thread 0 of every CTA stamps %globaltimer
    auto stamp = [&](int i) {
        if (trace && threadIdx.x == 0) {
            unsigned long long tns;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
            trace[(size_t)blockIdx.x * 8 + i] = tns;
        }
    };
    stamp(0);
    const int C = g.chunk, stage_bytes = C * BT;
    uint8_t* ring = smem + g.ring_off + (size_t)warp * g.stages * stage_bytes;
    const uint32_t bar0 = smem_u32(smem + g.bar_off) + warp * kMStagesMax * 8;

    // CTA c owns block-tiles [i0, i1); warp w the contiguous run [r0, r1) of them, streamed C at a time
    const int i0 = blockIdx.x * g.per_cta, i1 = min(g.total, i0 + g.per_cta);
    const int r0 = i0 + warp * g.per_warp, r1 = min(i1, r0 + g.per_warp);
    const int n_my = max(0, r1 - r0), n_steps = (n_my + C - 1) / C;

    if (lane == 0) {
        for (int s = 0; s < g.stages; s++) mbar_init(bar0 + s * 8, 1);
        fence_barrier_init();
    }
    __syncwarp();
    int issued = 0, ist = 0;
    auto issue_next = [&]() {
        if (lane == 0) {
            const uint32_t bar = bar0 + ist * 8;
            const uint32_t bytes = (uint32_t)min(C, n_my - issued * C) * BT;
            mbar_expect_tx(bar, bytes);
            bulk_g2s(smem_u32(ring + (size_t)ist * stage_bytes), wm + (size_t)(r0 + issued * C) * BT, bytes, bar);
        }
        issued++;
        if (++ist == g.stages) ist = 0;
    };
    const bool late = ms.sel != nullptr;   // expert weights are chosen by the previous kernel
    if (!late) {
        const int pre = min(n_steps, g.stages);
        for (int i = 0; i < pre; i++) issue_next();   // weights are constants: stream them before the dependency resolves
    }
    pdl_launch_dependents();
    stamp(1);
    pdl_wait();
    stamp(2);
    const float* ain = p.a;
    if (late) {
        const int ex = ms.sel[blockIdx.y];
        y += (size_t)blockIdx.y * ms.y_stride;
        if (ex < 0) {   // expert owned by another tensor-parallel rank: this slot contributes zeros to the all-reduce
            const int nout = pairs ? M / 2 : M;
            for (int i = blockIdx.x * kMT + threadIdx.x; i < nout; i += gridDim.x * kMT) y[i] = 0.0f;
            return;
        }
        wm += (size_t)ex * ms.stride;
        ain += (size_t)blockIdx.y * ms.a_stride;
        gpart += (size_t)blockIdx.y * ms.gpart_stride;
        const int pre = min(n_steps, g.stages);
        for (int i = 0; i < pre; i++) issue_next();
    }

    // ---- fused prologue, in registers: warp w builds super-blocks w, w+16, ... of x (zb_stream.cuh Prologue semantics:
    // v = a | silu(a)*a[K+i];  w1: v = rmsnorm(v, w1);  r: v += r, CTA 0 stores the residual stream;  w2: x = rmsnorm(v, w2))
    const int nb = g.nb;
    {
        const int nxb = (K + 255) >> 8;   // x is built in 256-element blocks: warp w owns blocks w, w+16, ...
        const int K4 = K >> 2;
        F8 xw[kMaxOwn];
#define xv(o) xw[o].v
        auto f4 = [&](int b, int h) { return TYPE == kQ6_K ? 64 * b + lane + 32 * h : 64 * b + 2 * lane + h; };
        const float* abase = ain + (p.a_rep > 1 ? (size_t)(blockIdx.x % p.a_rep) * p.a_rep_stride : 0);
        const float4* a4 = reinterpret_cast<const float4*>(abase);
        if (MIX) {   // MoE combine / K-slab partial sums: out = 0; out += y_k * w_k in order (moe.go:470-479)
#pragma unroll
            for (int o = 0; o < kMaxOwn; o++) {
                const int b = warp + o * kMW;
                if (b < nxb) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (f4(b, h) < K4) {
                            for (int k = 0; k < p.mix_n; k++) {
                                const float4 yk = __ldcg(reinterpret_cast<const float4*>(abase + (size_t)k * p.mix_stride) + f4(b, h));
                                const float wk = p.mix_w[k];
                                v.x = v.x + yk.x * wk; v.y = v.y + yk.y * wk; v.z = v.z + yk.z * wk; v.w = v.w + yk.w * wk;
                            }
                        }
                        xv(o)[4 * h] = v.x; xv(o)[4 * h + 1] = v.y; xv(o)[4 * h + 2] = v.z; xv(o)[4 * h + 3] = v.w;
                    }
                }
            }
        } else {   // the common case keeps its own straight-line loop: all loads of the warp's blocks issue back to back
#pragma unroll
            for (int o = 0; o < kMaxOwn; o++) {
                const int b = warp + o * kMW;
                if (b < nxb) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float4 v = f4(b, h) < K4 ? __ldcg(a4 + f4(b, h)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        xv(o)[4 * h] = v.x; xv(o)[4 * h + 1] = v.y; xv(o)[4 * h + 2] = v.z; xv(o)[4 * h + 3] = v.w;
                    }
                }
            }
        }
        if (p.swiglu) {   // silu_generic.go:22-31: sigmoid in f64, float32(g*sig)*u
            const float4* u4 = reinterpret_cast<const float4*>(abase + K);
#pragma unroll
            for (int o = 0; o < kMaxOwn; o++) {
                const int b = warp + o * kMW;
                if (b < nxb) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float4 u = f4(b, h) < K4 ? __ldcg(u4 + f4(b, h)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) xv(o)[4 * h + e] = silu_mul(xv(o)[4 * h + e], uu[e]);
                    }
                }
            }
        } else {
            auto sumsq = [&]() {
                float ss = 0.0f;
#pragma unroll
                for (int o = 0; o < kMaxOwn; o++)
                    if (warp + o * kMW < nxb) {
#pragma unroll
                        for (int e = 0; e < 8; e++) ss = fmaf(xv(o)[e], xv(o)[e], ss);
                    }
                return block_sum(ss, red);
            };
            auto scale_by = [&](float sc, const float* gain) {
                const float4* w4 = reinterpret_cast<const float4*>(gain);
#pragma unroll
                for (int o = 0; o < kMaxOwn; o++) {
                    const int b = warp + o * kMW;
                    if (b < nxb) {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const float4 w = f4(b, h) < K4 ? __ldg(w4 + f4(b, h)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            xv(o)[4 * h] = xv(o)[4 * h] * sc * w.x; xv(o)[4 * h + 1] = xv(o)[4 * h + 1] * sc * w.y;
                            xv(o)[4 * h + 2] = xv(o)[4 * h + 2] * sc * w.z; xv(o)[4 * h + 3] = xv(o)[4 * h + 3] * sc * w.w;
                        }
                    }
                }
            };
            if (p.w1) scale_by(inv_rms(sumsq(), K, p.eps), p.w1);
            if (p.r) {
                const float4* r4 = reinterpret_cast<const float4*>(p.r);
                float4* so4 = (blockIdx.x == 0 && blockIdx.y == 0 && p.sum_out) ? reinterpret_cast<float4*>(p.sum_out) : nullptr;
#pragma unroll
                for (int o = 0; o < kMaxOwn; o++) {
                    const int b = warp + o * kMW;
                    if (b < nxb) {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const float4 r = f4(b, h) < K4 ? __ldcg(r4 + f4(b, h)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            xv(o)[4 * h] += r.x; xv(o)[4 * h + 1] += r.y; xv(o)[4 * h + 2] += r.z; xv(o)[4 * h + 3] += r.w;
                            if (so4 && f4(b, h) < K4) so4[f4(b, h)] = make_float4(xv(o)[4 * h], xv(o)[4 * h + 1], xv(o)[4 * h + 2], xv(o)[4 * h + 3]);
                        }
                    }
                }
            }
            if (p.w2) scale_by(inv_rms(sumsq(), K, p.eps), p.w2);
        }
        stamp(3);
#pragma unroll
        for (int o = 0; o < kMaxOwn; o++) {
            const int b = warp + o * kMW;
            if (b < nxb) {
                if ((TYPE == kQ4_K || TYPE == kQ5_K) && I8) frags_q4k_i8(xw[o], b, lane, xf, xm, xinv);
                else if (TYPE == kQ4_K) frags_q4k(xw[o], b, lane, xf, xm, xinv);
                else if (TYPE == kQ6_K && I8) frags_q6k_i8(xw[o], b, lane, xf, xinv);
                else if (TYPE == kQ6_K) frags_q6k(xw[o], b, lane, reinterpret_cast<uint2*>(xf), reinterpret_cast<float*>(xm), xinv);
                else if (I8) frags_q40_i8(xw[o], b, lane, 64 * b + 2 * lane < K4, reinterpret_cast<uint2*>(xf), reinterpret_cast<int*>(xm), xinv);
                else frags_q40(xw[o], b, lane, 64 * b + 2 * lane < K4, reinterpret_cast<uint2*>(xf), reinterpret_cast<float*>(xm), xinv);
            }
        }
        if (TYPE != kQ4_K && TYPE != kQ5_K && threadIdx.x < 16) xm[nb * xm_stride(TYPE) + threadIdx.x] = 0u;   // the zero block lanes t != 0 read
#undef xv
    }
    __syncthreads();
    stamp(4);

    const int tau_first = i0 / nb;
    const int gq = lane >> 2, t = lane & 3;
    const int bsel = lane < 12 ? lane : (lane < 24 ? lane - 12 : lane - 24);   // lanes >= 12 duplicate fragments: their columns are ignored
    const uint32_t msel = (t & 1) ? 0x4342u : 0x4140u;
    // integer path: weights of the two digit columns this lane holds (columns 4..7 repeat digits and get weight zero)
    const float wlo = t == 0 ? 16777216.0f : (t == 1 ? 256.0f : 0.0f), whi = t == 0 ? 65536.0f : (t == 1 ? 1.0f : 0.0f);
    // Q6_K integer path: left-shift that brings the lane's int8 scale (byte is = t >> 1 of the even-qi pair, 2 + is of the odd) to the top
    const uint32_t q6selA = 24u - 8u * (uint32_t)(t >> 1), q6selB = 8u - 8u * (uint32_t)(t >> 1);
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    int tau = r0 / nb, b = r0 - tau * nb;
    int cur_tau = -1, st = 0;
    uint32_t parity = 0;
    auto flush = [&]() {   // this warp's share of row tile cur_tau -> its slot (warps that touch a tile are consecutive)
        float vlo = (I8 || t == 0) ? tot[0] + tot[1] : (t == 1 ? tot[0] : 0.0f);
        float vhi = (I8 || t == 0) ? tot[2] + tot[3] : (t == 1 ? tot[2] : 0.0f);
        vlo += __shfl_xor_sync(0xffffffffu, vlo, 1); vhi += __shfl_xor_sync(0xffffffffu, vhi, 1);
        vlo += __shfl_xor_sync(0xffffffffu, vlo, 2); vhi += __shfl_xor_sync(0xffffffffu, vhi, 2);
        if (t == 0) {
            const int w_first = (max(i0, cur_tau * nb) - i0) / g.per_warp;
            float* dst = part + (size_t)((cur_tau - tau_first) * g.slots + (warp - w_first)) * 16;
            dst[gq] = vlo;
            dst[gq + 8] = vhi;
        }
    };
    for (int j = 0; j < n_steps; j++) {
        const int cnt = min(C, n_my - j * C);
        mbar_wait(bar0 + st * 8, parity);
        for (int u = 0; u < cnt; u++) {
            if (tau != cur_tau) {
                if (cur_tau >= 0) flush();
                cur_tau = tau;
                tot[0] = tot[1] = tot[2] = tot[3] = 0.0f;
            }
            const uint8_t* bt = ring + (size_t)st * stage_bytes + (size_t)u * BT;
            if ((TYPE == kQ4_K || TYPE == kQ5_K) && I8)
                block_tile_q4k_i8<TYPE == kQ5_K>(bt, xf + (size_t)b * 96, xm + b * kXmWords, xinv[b], tot, lane, wlo, whi);
            else if (TYPE == kQ4_K)
                block_tile_q4k(bt, xf + (size_t)b * 96, xm + b * kXmWords, xinv[b], tot, lane, bsel, msel);
            else if (TYPE == kQ6_K && I8)
                block_tile_q6k_i8(bt, xf + (size_t)b * 96, xinv[b], tot, lane, q6selA, q6selB, (t & 1) ? 256.0f : 16777216.0f, (t & 1) ? 1.0f : 65536.0f);
            else if (TYPE == kQ6_K)
                block_tile_q6k(bt, xf + (size_t)b * 96, reinterpret_cast<const float*>(xm) + (t == 0 ? b : nb) * kXmWords, xinv[b], tot, lane, bsel);
            else if (I8)
                block_tile_q40_i8(bt, reinterpret_cast<const uint2*>(xf) + (size_t)b * 64, reinterpret_cast<const int*>(xm) + b * 16, xinv[b >> 1], tot,
                                  lane, wlo, whi);
            else
                block_tile_q40(bt, xf + (size_t)b * 48, reinterpret_cast<const float*>(xm) + (t == 0 ? b : nb) * 16, xinv[b >> 1], tot, lane, bsel);
            if (++b == nb) { b = 0; tau++; }
        }
        __syncwarp();
        if (issued < n_steps) {   // refill the stage just drained (generic-proxy reads ordered before the async-proxy write)
            fence_proxy_async();
            issue_next();
        }
        if (++st == g.stages) { st = 0; parity ^= 1u; }
    }
    if (cur_tau >= 0) flush();
    stamp(5);
    __syncthreads();
    stamp(6);

    // ---- per row tile: sum the warps' partials in slot order; a tile shared with other CTAs is finished by the owner of
    // its first part, which polls the (value, flag) pairs the others push (all CTAs are co-resident: grid <= SM count)
    const int n_local = i1 > i0 ? (i1 - 1) / nb - tau_first + 1 : 0;
    for (int base = 0; base < n_local * 16; base += kMT) {
        const int idx = base + threadIdx.x, tl = idx >> 4, row = idx & 15;
        const bool valid = tl < n_local;
        const int tt = tau_first + tl;
        const int lo = max(i0, tt * nb), hi = min(i1, (tt + 1) * nb);
        float v = 0.0f;
        if (valid) {
            const int ns = (hi - 1 - i0) / g.per_warp - (lo - i0) / g.per_warp + 1;
            for (int k = 0; k < ns; k++) v += part[(size_t)(tl * g.slots + k) * 16 + row];
        }
        const bool complete = (lo == tt * nb) && (hi == (tt + 1) * nb);
        bool fin = valid && complete;
        if (valid && !complete) {
            const int c_first = (tt * nb) / g.per_cta, c_last = ((tt + 1) * nb - 1) / g.per_cta;
            const int mypart = blockIdx.x - c_first;
            uint2* slot = gpart + ((size_t)tt * kMaxParts) * 16 + row;
            if (mypart != 0) {
                asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot + mypart * 16), "r"(__float_as_uint(v)), "r"(1u) : "memory");
            } else {
                for (int pp = 1; pp <= c_last - c_first; pp++) {
                    uint32_t val, flag, spins = 0;
                    do {
                        asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(val), "=r"(flag) : "l"(slot + pp * 16) : "memory");
                        if (++spins > (1u << 24)) __trap();   // a lost CTA traps instead of hanging the GPU
                    } while (flag != 1u);
                    v += __uint_as_float(val);
                    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot + pp * 16), "r"(0u), "r"(0u) : "memory");   // ready for the next launch
                }
                fin = true;
            }
        }
        __syncwarp();
        const float up = __shfl_down_sync(0xffffffffu, v, 1);
        if (fin) {
            const int grow = tt * 16 + row;
            if (pairs) {
                if (!(row & 1) && grow + 1 < M) y[grow >> 1] = silu_mul(v, up);
            } else if (grow < M) {
                y[grow] = v;
            }
        }
    }
    stamp(7);
}

}  // namespace

// ===========================================================================
// C ABI (include/zb200.h)
// ===========================================================================
ZB_API int zb_mma_check(int qtype, int rows, int cols) {
    MGeom g{};
    return make_mgeom(qtype, rows, cols, g) ? 0 : (int)cudaErrorInvalidConfiguration;
}

// Work split of one launch (host-side; tests/test_mma_layout.py checks its invariants on the CPU):
// out[0..11] = units per row, row tiles, block-tiles, block-tiles per CTA, CTAs, block-tiles per warp, tiles per ring stage,
// ring stages, partial-sum slots per row tile, row tiles per CTA (max), dynamic shared memory bytes, bytes per block-tile.
ZB_API int zb_mma_geometry(int qtype, int rows, int cols, int max_ctas, int* out) {
    MGeom g{};
    if (!out || max_ctas <= 0 || !make_mgeom(qtype, rows, cols, g, max_ctas)) return cudaErrorInvalidConfiguration;
    const int v[12] = {g.nb, g.n_tiles, g.total, g.per_cta, g.ctas, g.per_warp, g.chunk, g.stages, g.slots, g.max_local, g.smem_bytes, bt_bytes(qtype)};
    for (int i = 0; i < 12; i++) out[i] = v[i];
    return 0;
}

ZB_API int zb_mma_layout(int qtype, int rows, int cols, int64_t* weight_bytes, int64_t* scratch_bytes) {
    MGeom g{};
    if (!make_mgeom(qtype, rows, cols, g)) return cudaErrorInvalidConfiguration;
    if (weight_bytes) *weight_bytes = (int64_t)g.total * bt_bytes(qtype);
    if (scratch_bytes) *scratch_bytes = (int64_t)g.n_tiles * kMaxParts * 16 * 8;   // (value, flag) per row, part and row tile
    return 0;
}

// Host-side repack of raw GGUF blocks into block-tiles (pure byte moves; rows past `rows` are zero blocks).
ZB_API int zb_mma_repack_host(int qtype, const void* raw, int rows, int cols, void* out) {
    MGeom g{};
    if (!make_mgeom(qtype, rows, cols, g)) return cudaErrorInvalidConfiguration;
    const uint8_t* src = static_cast<const uint8_t*>(raw);
    uint8_t* dst = static_cast<uint8_t*>(out);
    const int BT = bt_bytes(qtype);
    // every byte of a full 16-row tile is written below; only the ragged last row tile needs zero fill (no whole-buffer memset:
    // on a 42 GB model the extra pass over fresh pages cost as much as the permutation itself)
    if (rows % 16) memset(dst + (size_t)(g.n_tiles - 1) * g.nb * BT, 0, (size_t)g.nb * BT);
    if (qtype == zb::kQ4_0) {
        const int nblk = cols / 32;
        zb::zb_parallel_for(g.n_tiles, 64, [&](int64_t tau) {
            for (int b = 0; b < g.nb; b++) {
                uint8_t* bt = dst + ((size_t)tau * g.nb + b) * BT;
                for (int h = 0; h < 2; h++)
                    for (int gq = 0; gq < 8; gq++) {
                        const int row = (int)tau * 16 + gq + 8 * h;
                        if (row >= rows) continue;
                        for (int bi = 0; bi < 4; bi++) {
                            const uint8_t* blk = src + ((size_t)row * nblk + b * 4 + bi) * 18;   // fp16 d, 16 nibble bytes
                            memcpy(bt + 1024 + gq * 16 + h * 8 + bi * 2, blk, 2);
                            for (int t = 0; t < 4; t++) memcpy(bt + ((h * 32 + gq * 4 + t) * 16) + bi * 4, blk + 2 + 4 * t, 4);
                        }
                    }
            }
        });
        return 0;
    }
    if (qtype == zb::kQ6_K) {
        zb::zb_parallel_for(g.n_tiles, 64, [&](int64_t tau) {
            for (int b = 0; b < g.nb; b++) {
                uint8_t* bt = dst + ((size_t)tau * g.nb + b) * BT;
                for (int h = 0; h < 2; h++)
                    for (int gq = 0; gq < 8; gq++) {
                        const int row = (int)tau * 16 + gq + 8 * h;
                        if (row >= rows) continue;
                        const uint8_t* blk = src + ((size_t)row * g.nb + b) * 210;   // ql[128] qh[64] sc[16] d
                        memcpy(bt + 3072 + (h * 8 + gq) * 16, blk + 192, 16);
                        memcpy(bt + 3328 + (h * 8 + gq) * 2, blk + 208, 2);
                        for (int t = 0; t < 4; t++) {
                            uint8_t* oa = bt + ((h * 3 + 0) * 32 + gq * 4 + t) * 16;
                            uint8_t* ob = bt + ((h * 3 + 1) * 32 + gq * 4 + t) * 16;
                            uint8_t* oh = bt + ((h * 3 + 2) * 32 + gq * 4 + t) * 16;
                            for (int hf = 0; hf < 2; hf++)
                                for (int is = 0; is < 2; is++) {
                                    const int c = hf * 2 + is, l = 16 * is + 4 * t;
                                    memcpy(oa + 4 * c, blk + 64 * hf + l, 4);
                                    memcpy(ob + 4 * c, blk + 64 * hf + 32 + l, 4);
                                    memcpy(oh + 4 * c, blk + 128 + 32 * hf + l, 4);
                                }
                        }
                    }
            }
        });
        return 0;
    }
    zb::zb_parallel_for(g.n_tiles, 64, [&](int64_t tau) {
        for (int b = 0; b < g.nb; b++) {
            uint8_t* bt = dst + ((size_t)tau * g.nb + b) * BT;
            for (int h = 0; h < 2; h++)
                for (int gq = 0; gq < 8; gq++) {
                    const int row = (int)tau * 16 + gq + 8 * h;
                    if (row >= rows) continue;
                    const uint8_t* blk = src + ((size_t)row * g.nb + b) * (qtype == zb::kQ5_K ? 176 : 144);
                    memcpy(bt + 2048 + (h * 8 + gq) * 16, blk, 16);   // d, dmin, scales[12]
                    if (qtype == zb::kQ5_K)   // qh[32] behind the nibbles: the lane's bytes 8t..8t+7
                        for (int t = 0; t < 4; t++) memcpy(bt + 2304 + (h * 32 + gq * 4 + t) * 8, blk + 144 + 8 * t, 8);
                    for (int pp = 0; pp < 2; pp++)
                        for (int t = 0; t < 4; t++) {
                            uint8_t* o = bt + ((h * 2 + pp) * 32 + gq * 4 + t) * 16;
                            memcpy(o, blk + 16 + 32 * (2 * pp) + 8 * t, 8);
                            memcpy(o + 8, blk + 16 + 32 * (2 * pp + 1) + 8 * t, 8);
                        }
                }
        }
    });
    return 0;
}

// Phase timeline for tuning: with ZB_MMA_TRACE=1 every launch (up to 64) records 8 %globaltimer stamps per CTA.
static unsigned long long* g_trace = nullptr;
static int g_trace_launches = 0;
constexpr int kTraceMax = 64, kTraceStride = 148 * 8;
ZB_API int zb_mma_trace_read(unsigned long long* out, int max_launches) {
    if (!g_trace) return 0;
    int n = g_trace_launches < max_launches ? g_trace_launches : max_launches;
    if (n > kTraceMax) n = kTraceMax;
    cudaMemcpy(out, g_trace, (size_t)n * kTraceStride * 8, cudaMemcpyDeviceToHost);
    return n;
}

ZB_API int zb_gemv_mma_f32(const zb_mma_weight* w, const zb_prologue* p, float* y, void* scratch, int flags, zb_stream_t stream) {
    if (!w || !p || !y || !scratch || !w->data) return cudaErrorInvalidValue;
    if (p->n_wait > 0) return cudaErrorInvalidValue;   // the fused TP exchange stays on the CUDA-core kernel
    if (p->mix_n > 0 && (!p->mix_w || p->swiglu || w->expert_sel)) return cudaErrorInvalidValue;
    MGeom g{};
    MSel ms{};
    int nsel = 1;
    if (w->expert_sel) {
        if (w->n_sel <= 0 || w->n_sel > 16 || (w->rows & 15) || p->sum_out) return cudaErrorInvalidValue;
        nsel = w->n_sel;
    }
    // slots run side by side: all CTAs of all slots must be co-resident (the partial-sum exchange polls a neighbour CTA)
    if (!make_mgeom(w->qtype, w->rows, w->cols, g, ZB_SMS / nsel)) return cudaErrorInvalidConfiguration;
    if (w->expert_sel) {
        ms.sel = w->expert_sel; ms.stride = w->expert_stride; ms.a_stride = p->a_slot_stride; ms.y_stride = w->y_slot_stride;
        ms.gpart_stride = (long long)g.n_tiles * kMaxParts * 16;
    }
    if (w->epilogue == 1 && (w->rows & 1)) return cudaErrorInvalidValue;
    // Per device: opt-in shared memory of every instantiation, and how many CTAs of this kernel the device can hold at once.
    // The partial-sum exchange of a row tile shared by two CTAs spins on its neighbour, so the whole grid (all slots of a
    // MoE / column-slab launch included) MUST be co-resident: the grid is sized from the device's own SM count and occupancy
    // (a MIG slice or a smaller part has fewer than 148 SMs), never from the compile-time constant alone.
    static DeviceOnce once;
    static int resident[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    dev &= 63;
    if (cudaError_t e = once.ensure(1, [&] {
            cudaError_t r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ4_K, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ4_K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ5_K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ6_K, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ6_K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ4_0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ4_0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ4_K, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ5_K, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ6_K, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            if (r == cudaSuccess) r = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ4_0, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
            int sms = 0, per_sm = 0;
            if (r == cudaSuccess) r = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (r == cudaSuccess) r = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gemv_mma_kernel<zb::kQ6_K, 1>, kMT, kMSmem - 2048);
            if (r == cudaSuccess) resident[dev] = sms * per_sm;
            return r;
        }))
        return e;
    if (resident[dev] < nsel) return cudaErrorCooperativeLaunchTooLarge;   // not even one CTA per slot fits: use the streamed kernel
    if (resident[dev] < ZB_SMS) {   // fewer SMs than the B200's 148: re-split the matrix over what is there
        if (!make_mgeom(w->qtype, w->rows, w->cols, g, resident[dev] / nsel)) return cudaErrorInvalidConfiguration;
        if (w->expert_sel) ms.gpart_stride = (long long)g.n_tiles * kMaxParts * 16;
    }
    if ((long long)g.ctas * nsel > resident[dev]) return cudaErrorCooperativeLaunchTooLarge;
    static const int want_trace = env_int("ZB_MMA_TRACE", 0);
    unsigned long long* trace = nullptr;
    if (want_trace) {
        if (!g_trace) {
            if (cudaMalloc(&g_trace, (size_t)kTraceMax * kTraceStride * 8) != cudaSuccess) return cudaErrorMemoryAllocation;
            cudaMemset(g_trace, 0, (size_t)kTraceMax * kTraceStride * 8);
        }
        if (g_trace_launches < kTraceMax) trace = g_trace + (size_t)(g_trace_launches++) * kTraceStride;
    }
    zb::Prologue pr{p->a, p->r, p->w1, p->w2, p->sum_out, p->mix_w, p->mix_n, p->mix_stride, p->eps, p->swiglu, p->a_replicas, p->a_replica_stride};
    uint2* gpart = static_cast<uint2*>(scratch);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g.ctas, nsel, 1);
    cfg.blockDim = dim3(kMT, 1, 1);
    cfg.dynamicSmemBytes = g.smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (flags & 1) ? 1 : 0;
    static const int use_i8 = env_int("ZB_MMA_I8", 1);   // integer (exact) tensor path for the K-quants; 0: f16 path
    if (p->mix_n > 0) {   // combine prologue: integer path only
        if (!use_i8) return cudaErrorInvalidValue;
#define ZB_MIX_LAUNCH(T) cudaLaunchKernelEx(&cfg, gemv_mma_kernel<T, 1, true>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y, \
                                            w->epilogue == 1 ? 1 : 0, gpart, trace, ms)
        switch (w->qtype) {
            case zb::kQ5_K: return ZB_MIX_LAUNCH(zb::kQ5_K);
            case zb::kQ4_0: return ZB_MIX_LAUNCH(zb::kQ4_0);
            case zb::kQ6_K: return ZB_MIX_LAUNCH(zb::kQ6_K);
            default: return ZB_MIX_LAUNCH(zb::kQ4_K);
        }
#undef ZB_MIX_LAUNCH
    }
    if (w->qtype == zb::kQ5_K)   // integer path only
        return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<zb::kQ5_K, 1>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y,
                                  w->epilogue == 1 ? 1 : 0, gpart, trace, ms);
    if (w->qtype == zb::kQ4_0 && use_i8)
        return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<zb::kQ4_0, 1>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y,
                                  w->epilogue == 1 ? 1 : 0, gpart, trace, ms);
    if (w->qtype == zb::kQ4_0)
        return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<zb::kQ4_0, 0>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y,
                                  w->epilogue == 1 ? 1 : 0, gpart, trace, ms);
    if (w->qtype == zb::kQ6_K && use_i8)
        return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<zb::kQ6_K, 1>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y,
                                  w->epilogue == 1 ? 1 : 0, gpart, trace, ms);
    if (w->qtype == zb::kQ6_K)
        return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<zb::kQ6_K, 0>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y,
                                  w->epilogue == 1 ? 1 : 0, gpart, trace, ms);
    if (use_i8)
        return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<zb::kQ4_K, 1>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y,
                                  w->epilogue == 1 ? 1 : 0, gpart, trace, ms);
    return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<zb::kQ4_K, 0>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y,
                              w->epilogue == 1 ? 1 : 0, gpart, trace, ms);
}
