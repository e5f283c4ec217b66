// Persistent whole-token decode kernel for batch-1 dense models (sm_100a) -- see zb_mega.cuh for the program format.
//
// Why: a Llama-3.2-3B decode step is ~160 dependent matrix-vector products of 3-30 MB each.  As separate launches (even
// PDL-chained inside one CUDA graph) every one of them pays launch + drain + first-tile latency, ~4 us against 1-4 us of
// HBM streaming (profiles/r01_mma_phase_trace_i8.txt).  Here the step is ONE cooperative launch:
//   * 148 CTAs x 16 warps stay resident; ops are separated by a grid barrier (one atomic + one acquire poll per CTA);
//   * each warp owns one TMA ring for the whole launch.  Its producer (lane 0) walks the stream table -- the block-tile
//     runs this warp owns in GEMV 0, 1, 2, ... -- and keeps the ring full with cp.async.bulk chunks, independent of which
//     op the CTA is executing: while the CTA waits at a barrier, loads x, builds digit fragments or exchanges partial sums,
//     the next ops' weights are already landing in shared memory (28 MB of ring chip-wide ~ one to two whole ops);
//   * the arithmetic of a GEMV is gemv_mma_kernel's (same work split, same summation order: bit-identical outputs);
//     the attention stage is decode_attn_item on warps 0-7, (KV head, split) items strided over the CTAs;
//   * the lm_head epilogue keeps a per-CTA argmax candidate, CTA 0 finishes the argmax and the position bookkeeping.
// Everything a CTA reads that another CTA wrote earlier in the launch is read through L2 (ld.global.cg / volatile).
#include "zb_mega.cuh"

namespace {

using namespace zb;

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of the (co-resident) grid arrive; thread 0 polls until `target` arrivals have been counted since reset.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int spins = 0;
        while ((int)(ld_acquire_u32(counter) - target) < 0) {
            if (++spins > (1u << 24)) __trap();   // a lost CTA traps instead of hanging the GPU
        }
        __threadfence();
    }
    __syncthreads();
}

template <int EPL, int REP>
__device__ __noinline__ void attn_item_call(const AttnArgs* p, int kvh, int split, int pos, uint8_t* smraw, uint32_t bar, uint32_t parity,
                                            int* s_last) {
    decode_attn_item<EPL, REP, kMegaAttnWarps, 1>(*p, kvh, split, 0, pos, smraw, bar, parity, s_last);
}

template <int EPL>
__device__ __forceinline__ void attn_item_rep(int rep, const AttnArgs* p, int kvh, int split, int pos, uint8_t* smraw, uint32_t bar,
                                              uint32_t parity, int* s_last) {
    switch (rep) {
        case 1: attn_item_call<EPL, 1>(p, kvh, split, pos, smraw, bar, parity, s_last); break;
        case 2: attn_item_call<EPL, 2>(p, kvh, split, pos, smraw, bar, parity, s_last); break;
        case 3: attn_item_call<EPL, 3>(p, kvh, split, pos, smraw, bar, parity, s_last); break;
        case 4: attn_item_call<EPL, 4>(p, kvh, split, pos, smraw, bar, parity, s_last); break;
        case 8: attn_item_call<EPL, 8>(p, kvh, split, pos, smraw, bar, parity, s_last); break;
    }
}

// phase timeline for tuning: thread 0 of every CTA stamps the SM clock
__device__ __forceinline__ void mega_stamp(long long* trace, int op, int k) {
    if (trace && threadIdx.x == 0) trace[((size_t)op * gridDim.x + blockIdx.x) * kMegaTraceSlots + k] = clock64();
}

__device__ __forceinline__ float softcap_apply(float v, float cap, float inv_cap) {   // arch_llama.go:15-27,184-213
    float x = v * inv_cap, t;
    if (x > 4.5f) t = 1.0f;
    else if (x < -4.5f) t = -1.0f;
    else {
        float x2 = x * x;
        t = x * (27.0f + x2) / (27.0f + 9.0f * x2);
    }
    return cap * t;
}

// ---- per-warp TMA ring: a byte FIFO of chunks (1..4 block-tiles) in the order the warp will consume them -----------------
// State lives in shared memory (the kernel has no registers to spare): all lanes read it, lane 0 updates it, __syncwarp orders.
struct WarpRing {
    int q_off[kMegaRingBars];      // ring offset of chunk (seq % kMegaRingBars)
    int ps, pj, pn, p_nmy, p_bt, p_C;   // producer cursor: stream, chunk in the warp's run, chunks / block-tiles of the run, tile bytes, tiles per chunk
    unsigned long long p_src;      // first byte of the run
    int head, inflight;            // next write offset, chunks issued and not yet consumed
    unsigned int p_seq, c_seq;     // chunks issued / consumed since launch
};

struct MegaShared {
    MegaOp op;
    float red[32];
    WarpRing ring[kMW];
    unsigned long long attn_bar;
    int s_last;
    float cv[kMW];
    int ci[kMW];
};

// lane 0: move the producer cursor to the next stream in which this warp owns block-tiles
__device__ __forceinline__ void ring_advance(volatile WarpRing& r, const MegaStream* __restrict__ streams, int n_act, int warp) {
    int ps = r.ps;
    while (++ps < n_act) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(streams + ps));
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(streams + ps) + 1);
        const int total = (int)lo.z, per_cta = (int)lo.w, per_warp = (int)hi.x;
        const int i0 = blockIdx.x * per_cta, i1 = min(total, i0 + per_cta);
        const int r0 = i0 + warp * per_warp, r1 = min(i1, r0 + per_warp);
        if (r1 > r0) {
            const int bt = (int)hi.y, C = (int)hi.z;
            r.p_nmy = r1 - r0;
            r.p_bt = bt;
            r.p_C = C;
            r.pn = (r1 - r0 + C - 1) / C;
            r.pj = 0;
            r.p_src = (((unsigned long long)lo.y << 32) | lo.x) + (unsigned long long)r0 * (unsigned long long)bt;
            break;
        }
    }
    r.ps = ps;
}

// All lanes of the warp, uniformly: issue the next chunk if it fits.  Returns whether a chunk was issued.
__device__ __forceinline__ bool ring_try_issue(volatile WarpRing& r, const MegaStream* __restrict__ streams, int n_act, uint8_t* ring,
                                               uint32_t bar0, int ring_w, int warp, int lane) {
    const int ps = r.ps, inflight = r.inflight;
    if (ps >= n_act || inflight >= kMegaRingBars) return false;
    const int pj = r.pj, C = r.p_C, bt = r.p_bt;
    const int cnt = min(C, r.p_nmy - pj * C), bytes = cnt * bt;
    const int head = r.head;
    int at;
    if (inflight == 0) {
        at = 0;
    } else {
        const int tail = r.q_off[r.c_seq & (kMegaRingBars - 1)];   // oldest chunk not yet consumed
        if (head > tail) {
            if (head + bytes <= ring_w) at = head;
            else if (bytes <= tail) at = 0;
            else return false;
        } else if (head < tail) {
            if (head + bytes <= tail) at = head;
            else return false;
        } else {
            return false;   // full
        }
    }
    __syncwarp();   // every lane has read the state lane 0 is about to change
    if (lane == 0) {
        const unsigned int seq = r.p_seq;
        const int slot = seq & (kMegaRingBars - 1);
        r.q_off[slot] = at;
        const uint32_t bar = bar0 + slot * 8;
        mbar_expect_tx(bar, (uint32_t)bytes);
        bulk_g2s(smem_u32(ring + at), reinterpret_cast<const uint8_t*>(r.p_src) + (size_t)pj * C * bt, (uint32_t)bytes, bar);
        r.head = at + bytes;
        r.inflight = inflight + 1;
        r.p_seq = seq + 1;
        r.pj = pj + 1;
        if (pj + 1 == r.pn) ring_advance(r, streams, n_act, warp);
    }
    __syncwarp();
    return true;
}

// ---- fused activation prologue of a GEMV, entirely in registers (gemv_mma_kernel's, as always-inlined helpers: a lambda that is
// not inlined would take the address of the register arrays and push them to local memory) ---------------------------------
template <int TYPE>
__device__ __forceinline__ int mega_f4(int b, int h, int lane) { return TYPE == kQ6_K ? 64 * b + lane + 32 * h : 64 * b + 2 * lane + h; }

template <int TYPE>
__device__ __forceinline__ float mega_sumsq(const F8 (&xw)[kMaxOwn], int nxb, int warp, float* red) {
    float ss = 0.0f;
#pragma unroll
    for (int o = 0; o < kMaxOwn; o++)
        if (warp + o * kMW < nxb) {
#pragma unroll
            for (int e = 0; e < 8; e++) ss = fmaf(xw[o].v[e], xw[o].v[e], ss);
        }
    return block_sum(ss, red);
}

template <int TYPE>
__device__ __forceinline__ void mega_scale_by(F8 (&xw)[kMaxOwn], float sc, const float* gain, int nxb, int K4, int warp, int lane) {
    const float4* w4 = reinterpret_cast<const float4*>(gain);
#pragma unroll
    for (int o = 0; o < kMaxOwn; o++) {
        const int b = warp + o * kMW;
        if (b < nxb) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int i = mega_f4<TYPE>(b, h, lane);
                const float4 w = i < K4 ? __ldg(w4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                xw[o].v[4 * h] = xw[o].v[4 * h] * sc * w.x; xw[o].v[4 * h + 1] = xw[o].v[4 * h + 1] * sc * w.y;
                xw[o].v[4 * h + 2] = xw[o].v[4 * h + 2] * sc * w.z; xw[o].v[4 * h + 3] = xw[o].v[4 * h + 3] * sc * w.w;
            }
        }
    }
}

template <int TYPE>
__device__ __forceinline__ void mega_build_frags(const Prologue& p, int K, uint4* xf, uint32_t* xm, float* xinv, float* red, int warp, int lane) {
    const int nxb = (K + 255) >> 8;
    const int K4 = K >> 2;
    F8 xw[kMaxOwn];
    const float4* a4 = reinterpret_cast<const float4*>(p.a);
#pragma unroll
    for (int o = 0; o < kMaxOwn; o++) {
        const int b = warp + o * kMW;
        if (b < nxb) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int i = mega_f4<TYPE>(b, h, lane);
                const float4 v = i < K4 ? __ldcg(a4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                xw[o].v[4 * h] = v.x; xw[o].v[4 * h + 1] = v.y; xw[o].v[4 * h + 2] = v.z; xw[o].v[4 * h + 3] = v.w;
            }
        }
    }
    if (p.swiglu) {   // silu_generic.go:22-31: sigmoid in f64, float32(g*sig)*u
        const float4* u4 = reinterpret_cast<const float4*>(p.a + K);
#pragma unroll
        for (int o = 0; o < kMaxOwn; o++) {
            const int b = warp + o * kMW;
            if (b < nxb) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = mega_f4<TYPE>(b, h, lane);
                    const float4 u = i < K4 ? __ldcg(u4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    xw[o].v[4 * h] = silu_mul(xw[o].v[4 * h], u.x); xw[o].v[4 * h + 1] = silu_mul(xw[o].v[4 * h + 1], u.y);
                    xw[o].v[4 * h + 2] = silu_mul(xw[o].v[4 * h + 2], u.z); xw[o].v[4 * h + 3] = silu_mul(xw[o].v[4 * h + 3], u.w);
                }
            }
        }
    } else {
        if (p.w1) mega_scale_by<TYPE>(xw, inv_rms(mega_sumsq<TYPE>(xw, nxb, warp, red), K, p.eps), p.w1, nxb, K4, warp, lane);
        if (p.r) {
            const float4* r4 = reinterpret_cast<const float4*>(p.r);
            float4* so4 = (blockIdx.x == 0 && p.sum_out) ? reinterpret_cast<float4*>(p.sum_out) : nullptr;
#pragma unroll
            for (int o = 0; o < kMaxOwn; o++) {
                const int b = warp + o * kMW;
                if (b < nxb) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int i = mega_f4<TYPE>(b, h, lane);
                        const float4 r = i < K4 ? __ldcg(r4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                        xw[o].v[4 * h] += r.x; xw[o].v[4 * h + 1] += r.y; xw[o].v[4 * h + 2] += r.z; xw[o].v[4 * h + 3] += r.w;
                        if (so4 && i < K4) so4[i] = make_float4(xw[o].v[4 * h], xw[o].v[4 * h + 1], xw[o].v[4 * h + 2], xw[o].v[4 * h + 3]);
                    }
                }
            }
        }
        if (p.w2) mega_scale_by<TYPE>(xw, inv_rms(mega_sumsq<TYPE>(xw, nxb, warp, red), K, p.eps), p.w2, nxb, K4, warp, lane);
    }
#pragma unroll
    for (int o = 0; o < kMaxOwn; o++) {
        const int b = warp + o * kMW;
        if (b < nxb) {
            if (TYPE == kQ4_K || TYPE == kQ5_K) frags_q4k_i8(xw[o], b, lane, xf, xm, xinv);
            else if (TYPE == kQ6_K) frags_q6k_i8(xw[o], b, lane, xf, xinv);
            else frags_q40_i8(xw[o], b, lane, 64 * b + 2 * lane < K4, reinterpret_cast<uint2*>(xf), reinterpret_cast<int*>(xm), xinv);
        }
    }
}

// this warp's share of row tile `tl` (CTA-local index) -> its slot; warps that touch a tile are consecutive
__device__ __forceinline__ void mega_flush(const float (&tot)[4], float* part, int tl, int slots, int wslot, int lane) {
    float vlo = tot[0] + tot[1], vhi = tot[2] + tot[3];
    vlo += __shfl_xor_sync(0xffffffffu, vlo, 1); vhi += __shfl_xor_sync(0xffffffffu, vhi, 1);
    vlo += __shfl_xor_sync(0xffffffffu, vlo, 2); vhi += __shfl_xor_sync(0xffffffffu, vhi, 2);
    if ((lane & 3) == 0) {
        float* dst = part + (size_t)(tl * slots + wslot) * 16;
        dst[lane >> 2] = vlo;
        dst[(lane >> 2) + 8] = vhi;
    }
}

// ---- one GEMV op: fused prologue in registers, the warp's run of block-tiles from its ring, partial-sum exchange, epilogue -----
template <int TYPE>
__device__ __forceinline__ void gemv_phase(const MegaCtl& c, MegaShared* sh, uint8_t* smem, int with_head, int oi) {
    // The op descriptor lives in shared memory: every stage below re-reads what it needs after the barrier that precedes it,
    // so that nothing but the stage's own working set is live in registers (the Q6_K tile alone takes ~120).
    const MegaGemv& g = sh->op.g;
    float* red = sh->red;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BT = bt_bytes(TYPE);
    if (blockIdx.x * g.per_cta >= g.total) {   // CTA-uniform: no block-tiles of this matrix here
        if (g.head && threadIdx.x == 0) {
            c.cand_v[blockIdx.x] = -FLT_MAX;
            c.cand_i[blockIdx.x] = 0x7fffffff;
        }
        return;
    }

    // ---- fused prologue, in registers: warp w builds super-blocks w, w+16, ... of x (zb_stream.cuh Prologue semantics)
    mega_build_frags<TYPE>(g.p, g.K, reinterpret_cast<uint4*>(smem + g.xf_off), reinterpret_cast<uint32_t*>(smem + g.xm_off),
                           reinterpret_cast<float*>(smem + g.xinv_off), red, warp, lane);
    __syncthreads();
    mega_stamp(c.trace, oi, 1);

    // ---- main loop: this warp's run of block-tiles, chunk by chunk from its ring (gemv_mma_kernel's arithmetic and order)
    const int nb = g.nb;
    const int i0 = blockIdx.x * g.per_cta, i1 = min(g.total, i0 + g.per_cta);
    float* part = reinterpret_cast<float*>(smem + g.part_off);
    volatile WarpRing& rs = sh->ring[warp];
    uint8_t* const ring = smem + c.region_bytes + (size_t)warp * c.ring_w;
    const uint32_t bar0 = smem_u32(smem + c.region_bytes + (size_t)kMW * c.ring_w) + warp * kMegaRingBars * 8;
    const int n_act = with_head ? c.n_streams : c.n_streams_nohead;
    const int tau_first = i0 / nb;
    const int t = lane & 3;
    {
        const int C = g.chunk;
        const int r0 = i0 + warp * g.per_warp, r1 = min(i1, r0 + g.per_warp);
        const int n_my = max(0, r1 - r0), n_steps = (n_my + C - 1) / C;
        const uint4* xf = reinterpret_cast<const uint4*>(smem + g.xf_off);
        const uint32_t* xm = reinterpret_cast<const uint32_t*>(smem + g.xm_off);
        const float* xinv = reinterpret_cast<const float*>(smem + g.xinv_off);
        const float wlo = t == 0 ? 16777216.0f : (t == 1 ? 256.0f : 0.0f), whi = t == 0 ? 65536.0f : (t == 1 ? 1.0f : 0.0f);
        const uint32_t q6selA = 24u - 8u * (uint32_t)(t >> 1), q6selB = 8u - 8u * (uint32_t)(t >> 1);
        float tot[4] = {0.f, 0.f, 0.f, 0.f};
        int tau = r0 / nb, b = r0 - tau * nb;
        int cur_tau = -1;
        unsigned int cs = rs.c_seq;
        for (int j = 0; j < n_steps; j++) {
            const int cnt = min(C, n_my - j * C);
            const int slot = cs & (kMegaRingBars - 1);
            mbar_wait(bar0 + slot * 8, (cs / kMegaRingBars) & 1u);
            const uint8_t* chunk = ring + rs.q_off[slot];
            for (int u = 0; u < cnt; u++) {
                if (tau != cur_tau) {
                    if (cur_tau >= 0) mega_flush(tot, part, cur_tau - tau_first, g.slots, warp - (max(i0, cur_tau * nb) - i0) / g.per_warp, lane);
                    cur_tau = tau;
                    tot[0] = tot[1] = tot[2] = tot[3] = 0.0f;
                }
                const uint8_t* bt = chunk + (size_t)u * BT;
                if (TYPE == kQ4_K || TYPE == kQ5_K)
                    block_tile_q4k_i8<TYPE == kQ5_K>(bt, xf + (size_t)b * 96, xm + b * kXmWords, xinv[b], tot, lane, wlo, whi);
                else if (TYPE == kQ6_K)
                    block_tile_q6k_i8(bt, xf + (size_t)b * 96, xinv[b], tot, lane, q6selA, q6selB, (t & 1) ? 256.0f : 16777216.0f,
                                      (t & 1) ? 1.0f : 65536.0f);
                else
                    block_tile_q40_i8(bt, reinterpret_cast<const uint2*>(xf) + (size_t)b * 64, reinterpret_cast<const int*>(xm) + b * 16,
                                      xinv[b >> 1], tot, lane, wlo, whi);
                if (++b == nb) { b = 0; tau++; }
            }
            __syncwarp();
            fence_proxy_async();   // generic-proxy reads of the chunk ordered before the async-proxy refill of its bytes
            cs++;
            if (lane == 0) {
                rs.c_seq = cs;
                rs.inflight = rs.inflight - 1;
            }
            __syncwarp();
            while (ring_try_issue(rs, c.streams, n_act, ring, bar0, c.ring_w, warp, lane)) {}
        }
        if (cur_tau >= 0) mega_flush(tot, part, cur_tau - tau_first, g.slots, warp - (max(i0, cur_tau * nb) - i0) / g.per_warp, lane);
    }
    mega_stamp(c.trace, oi, 2);
    __syncthreads();
    mega_stamp(c.trace, oi, 3);

    // ---- per row tile: sum the warps' partials in slot order; a tile shared with other CTAs is finished by the owner of
    // its first part, which polls the (value, flag) pairs the others push (all CTAs are co-resident: cooperative launch)
    uint2* gpart = c.gpart + (size_t)g.region * c.gpart_stride;
    const int M = g.M;
    const float cap = g.softcap, inv_cap = cap > 0.0f ? (float)(1.0 / (double)cap) : 0.0f;
    float best_v = -FLT_MAX;
    int best_i = 0x7fffffff;
    const int n_local = (i1 - 1) / nb - tau_first + 1;
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
                    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot + pp * 16), "r"(0u), "r"(0u) : "memory");   // ready for reuse
                }
                fin = true;
            }
        }
        __syncwarp();
        const float up = __shfl_down_sync(0xffffffffu, v, 1);
        if (fin) {
            const int grow = tt * 16 + row;
            if (g.pairs) {
                if (!(row & 1) && grow + 1 < M) g.y[grow >> 1] = silu_mul(v, up);
            } else if (grow < M) {
                if (cap > 0.0f) v = softcap_apply(v, cap, inv_cap);
                g.y[grow] = v;
                if (g.head && (v > best_v || (v == best_v && grow < best_i))) { best_v = v; best_i = grow; }
            }
        }
    }
    if (g.head) {   // this CTA's argmax candidate: highest value, lowest index among equals (argmax.cu:40-48)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            const int oi2 = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best_v || (ov == best_v && oi2 < best_i)) { best_v = ov; best_i = oi2; }
        }
        if (lane == 0) { sh->cv[warp] = best_v; sh->ci[warp] = best_i; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kMW; w++)
                if (sh->cv[w] > best_v || (sh->cv[w] == best_v && sh->ci[w] < best_i)) { best_v = sh->cv[w]; best_i = sh->ci[w]; }
            c.cand_v[blockIdx.x] = best_v;
            c.cand_i[blockIdx.x] = best_i;
        }
    }
}

__device__ __noinline__ void attn_phase(MegaShared* sh, uint8_t* smem, uint32_t& attn_parity) {
    const AttnArgs& a = sh->op.a;
    const int warp = threadIdx.x >> 5;
    const int G = gridDim.x;
    const uint32_t attn_bar = smem_u32(&sh->attn_bar);
    const int pos = *a.pos_ptr;
    const int len = pos + 1, nsplits = (len + a.chunk - 1) / a.chunk, items = a.nkv * nsplits;
    const int rep = a.nq / a.nkv;
    uint32_t par = attn_parity;
    for (int it = blockIdx.x; it < items; it += G) {
        const int kvh = it / nsplits, split = it - kvh * nsplits;
        if (warp < kMegaAttnWarps) {
            switch (a.hd) {
                case 32: attn_item_rep<1>(rep, &a, kvh, split, pos, smem, attn_bar, par, &sh->s_last); break;
                case 64: attn_item_rep<2>(rep, &a, kvh, split, pos, smem, attn_bar, par, &sh->s_last); break;
                case 128: attn_item_rep<4>(rep, &a, kvh, split, pos, smem, attn_bar, par, &sh->s_last); break;
                case 256: attn_item_rep<8>(rep, &a, kvh, split, pos, smem, attn_bar, par, &sh->s_last); break;
            }
        }
        par ^= 1u;
    }
    attn_parity = par;
}

__device__ __noinline__ void embed_phase(MegaShared* sh) {
    // token select + embedding row gather, bit-exact dequantisation (+ Gemma scale): arch_llama.go:246-342, arch_gemma.go:38
    const MegaEmbed& e = sh->op.e;
    const int fi = *e.feed_idx;
    int tok = fi < *e.feed_len ? e.feed[fi] : *e.last;
    if (tok < 0) tok = 0;
    if (tok >= e.vocab) tok = e.vocab - 1;
    const int64_t base = (int64_t)tok * e.hidden;
    for (int i = blockIdx.x * kMT + threadIdx.x; i < e.hidden; i += gridDim.x * kMT) {
        const float v = deq_raw(e.type, e.table, base + i);
        e.out[i] = e.scale > 0.0f ? v * e.scale : v;
    }
}

__device__ __noinline__ void final_phase(const MegaCtl* cp, MegaShared* sh, int with_head) {
    // argmax over the CTAs' candidates + step bookkeeping (sampling_helpers.go:11-45, tensor_cache.go:205-262 counters)
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const MegaCtl& c = *cp;
    const MegaFinal& f = sh->op.f;
    const int lane = threadIdx.x, G = gridDim.x;
    int tok = 0;
    if (with_head) {
        float bv = -FLT_MAX;
        int bi = 0x7fffffff;
        for (int i = lane; i < G; i += 32) {
            const float v = __ldcg(c.cand_v + i);
            const int ix = __ldcg(c.cand_i + i);
            if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi2 = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi2 < bi)) { bv = ov; bi = oi2; }
        }
        tok = bi;
    }
    if (lane == 0) {
        *f.pos += 1;
        *f.step += 1;
        if (*f.feed_idx < *f.feed_len) *f.feed_idx += 1;
        if (with_head) {
            *f.amax = tok;
            *f.last = tok;
            const int n = *f.n_out;
            if (n < f.out_cap) f.out[n] = tok;
            *f.n_out = n + 1;
        }
    }
}

__global__ void __launch_bounds__(kMT, 1) decode_mega_kernel(const __grid_constant__ MegaCtl c, int with_head) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(16) MegaShared sh;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;
    {
        const uint32_t bar0 = smem_u32(smem + c.region_bytes + (size_t)kMW * c.ring_w) + warp * kMegaRingBars * 8;
        if (lane == 0) {
            for (int s = 0; s < kMegaRingBars; s++) mbar_init(bar0 + s * 8, 1);
            if (warp == 0) mbar_init(smem_u32(&sh.attn_bar), 1);
            fence_barrier_init();
            volatile WarpRing& r = sh.ring[warp];
            r.ps = -1; r.pj = 0; r.pn = 0; r.p_nmy = 0; r.p_bt = 0; r.p_C = 1; r.p_src = 0ull;
            r.head = 0; r.inflight = 0; r.p_seq = 0u; r.c_seq = 0u;
            ring_advance(r, c.streams, with_head ? c.n_streams : c.n_streams_nohead, warp);
        }
        __syncthreads();
        // first fill: the weights are constants, stream them before anything else happens
        while (ring_try_issue(sh.ring[warp], c.streams, with_head ? c.n_streams : c.n_streams_nohead,
                              smem + c.region_bytes + (size_t)warp * c.ring_w, bar0, c.ring_w, warp, lane)) {}
    }

    const unsigned int bar_base = (unsigned int)__ldcg(c.step) * (unsigned int)c.n_barriers * (unsigned int)G;
    unsigned int bar_k = 0;
    uint32_t attn_parity = 0;

    for (int oi = 0; oi < c.n_ops; oi++) {
        __syncthreads();   // everybody is done with the previous op's descriptor and scratch region
        if (threadIdx.x < (int)(sizeof(MegaOp) / 4))
            reinterpret_cast<uint32_t*>(&sh.op)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(c.ops + oi) + threadIdx.x);
        __syncthreads();
        const int kind = sh.op.kind;
        mega_stamp(c.trace, oi, 0);
        if (kind == kMegaGemv) {
            if (!(sh.op.g.head && !with_head)) {
                switch (sh.op.g.type) {
                    case kQ4_K: gemv_phase<kQ4_K>(c, &sh, smem, with_head, oi); break;
                    case kQ5_K: gemv_phase<kQ5_K>(c, &sh, smem, with_head, oi); break;
                    case kQ6_K: gemv_phase<kQ6_K>(c, &sh, smem, with_head, oi); break;
                    default: gemv_phase<kQ4_0>(c, &sh, smem, with_head, oi); break;
                }
            }
        } else if (kind == kMegaAttn) {
            attn_phase(&sh, smem, attn_parity);
        } else if (kind == kMegaEmbed) {
            embed_phase(&sh);
        } else {
            final_phase(&c, &sh, with_head);
        }
        mega_stamp(c.trace, oi, 4);
        if (sh.op.barrier) {
            bar_k++;
            grid_barrier(c.bar_counter, bar_base + bar_k * (unsigned int)G);
        }
        mega_stamp(c.trace, oi, 5);
    }
}

}  // namespace

namespace zb {

bool mega_attn_supported(int head_dim, int rep) {
    const bool hd_ok = head_dim == 32 || head_dim == 64 || head_dim == 128 || head_dim == 256;
    const bool rep_ok = rep == 1 || rep == 2 || rep == 3 || rep == 4 || rep == 8;
    return hd_ok && rep_ok && kMegaAttnWarps * rep <= 64;   // the per-warp partial outputs reuse the 32-position K/V tile
}

static int mega_configure(int device) {
    static bool done[64] = {false};
    if (device >= 0 && device < 64 && done[device]) return 0;
    cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMegaSmem);
    if (e != cudaSuccess) return (int)e;
    if (device >= 0 && device < 64) done[device] = true;
    return 0;
}

// One CTA per SM, all co-resident (the grid barrier and the partial-sum exchange spin on other CTAs).
int mega_max_ctas(int device, int* out_ctas) {
    if (int rc = mega_configure(device)) return rc;
    int sms = 0, per_sm = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_mega_kernel, kMT, kMegaSmem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
    *out_ctas = sms < ZB_SMS ? sms : ZB_SMS;
    return 0;
}

int mega_launch(const MegaCtl& ctl, int ctas, int with_head, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas, 1, 1);
    cfg.blockDim = dim3(kMT, 1, 1);
    cfg.dynamicSmemBytes = kMegaSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // the driver refuses a grid that cannot be co-resident
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, decode_mega_kernel, ctl, with_head);
}

}  // namespace zb
