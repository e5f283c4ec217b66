// Persistent whole-token decode kernel for batch-1 dense models (sm_100a) -- see zb_mega.cuh for the program format.
//
// Why: a Llama-3.2-3B decode step is ~160 dependent matrix-vector products of 3-30 MB each.  As separate launches (even
// PDL-chained inside one CUDA graph) every one of them pays launch + drain + first-tile latency, ~4 us against 1-4 us of
// HBM streaming (profiles/r01_mma_phase_trace_i8.txt).  Here the step is ONE cooperative launch:
//   * 148 CTAs x (16 consumer warps + 1 producer warp) stay resident; an op hands its output to the next as flagged
//     (value, epoch) pairs that the readers poll (zb_mega.cuh MegaVec) -- no grid barrier between GEMVs, row tiles that
//     straddle two CTAs are published as partial-sum planes the consumer adds; one grid barrier per launch (before the argmax);
//   * the producer warp walks the stream table -- the block-tiles this CTA owns in GEMV 0, 1, 2, ... in the round-robin order
//     its consumer warps will want them -- and keeps ONE CTA-wide ring of block-tile slots full with cp.async.bulk copies
//     (full / empty mbarrier per slot), independent of which op the consumers are executing: while they wait at a grid
//     barrier, load x, build digit fragments or exchange partial sums, the next ops' weights are already landing in shared
//     memory (~20 MB of ring chip-wide, more than one whole op).  The weight stream is tagged L2 evict-first, so the small
//     vectors every CTA re-reads (activations, norm gains, the op table) stay L2-resident across steps;
//   * the arithmetic of a GEMV is gemv_mma_kernel's (same work split, same summation order);
//     the attention stage is decode_attn_item on warps 0-7, (KV head, split) items strided over the CTAs;
//   * the lm_head epilogue keeps a per-CTA argmax candidate, CTA 0 finishes the argmax and the position bookkeeping.
// Everything a CTA reads that another CTA wrote earlier in the launch is read through L2 (ld.volatile / ld.global.cg).
// Measured slower than the CUDA-graph step on C2 (1.61 vs 1.21 ms/step): DESIGN.md 4.6 has the phase analysis.
#include "zb_mega.cuh"

namespace {

using namespace zb;

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The 16 consumer warps synchronise on a named barrier: the producer warp runs ahead on its own and never joins.
__device__ __forceinline__ void csync() { asm volatile("bar.sync 2, %0;" ::"n"(kMT) : "memory"); }

// consumer-wide sum (16 warps); `red` holds 32 floats
__device__ __forceinline__ float csum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    csync();   // protect `red` against a previous use
    if (l == 0) red[w] = v;
    csync();
    return warp_sum(l < kMW ? red[l] : 0.0f);
}

// All CTAs of the (co-resident) grid arrive; thread 0 polls until `target` arrivals have been counted since reset.
// Arrival is a fire-and-forget release reduction (no round trip for a return value); bar.sync makes the CTA's earlier
// writes part of what thread 0 releases, and the acquire poll + bar.sync hands the other CTAs' writes to every thread.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
    csync();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned int spins = 0;
        while ((int)(ld_acquire_u32(counter) - target) < 0) {
            if (++spins > (1u << 24)) __trap();   // a lost CTA traps instead of hanging the GPU
        }
    }
    csync();
}

// Pure spin (try_wait suspends the thread for a bounded time by itself): the back-off sleep of zb::mbar_wait would add its
// wake-up latency to every ring hand-over here.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// weights are read exactly once per token: evict-first keeps them from flushing the vectors every CTA re-reads out of L2
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}

template <int EPL, int REP>
__device__ __noinline__ void attn_item_call(const AttnArgs* p, int kvh, int split, int pos, uint8_t* smraw, uint32_t bar, uint32_t parity,
                                            int* s_last, uint32_t epoch_in, uint32_t epoch_out) {
    decode_attn_item<EPL, REP, kMegaAttnWarps, 1>(*p, kvh, split, 0, pos, smraw, bar, parity, s_last, epoch_in, epoch_out);
}

template <int EPL>
__device__ __forceinline__ void attn_item_rep(int rep, const AttnArgs* p, int kvh, int split, int pos, uint8_t* smraw, uint32_t bar,
                                              uint32_t parity, int* s_last, uint32_t ei, uint32_t eo) {
    switch (rep) {
        case 1: attn_item_call<EPL, 1>(p, kvh, split, pos, smraw, bar, parity, s_last, ei, eo); break;
        case 2: attn_item_call<EPL, 2>(p, kvh, split, pos, smraw, bar, parity, s_last, ei, eo); break;
        case 3: attn_item_call<EPL, 3>(p, kvh, split, pos, smraw, bar, parity, s_last, ei, eo); break;
        case 4: attn_item_call<EPL, 4>(p, kvh, split, pos, smraw, bar, parity, s_last, ei, eo); break;
        case 8: attn_item_call<EPL, 8>(p, kvh, split, pos, smraw, bar, parity, s_last, ei, eo); break;
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

// ---- CTA-wide TMA ring of block-tile slots ---------------------------------------------------------------------------
// Sequence number of a block-tile inside this CTA: GEMVs in op order; inside one GEMV round j = 0, 1, ... of the consumer
// warps' runs, warp by warp (warps 0..W_full-1 own per_warp tiles, warp W_full the remainder n_part).  Slot = seq % nslots,
// mbarrier phase = (seq / nslots) & 1.  Both sides compute the same numbers; nothing else is shared.
struct MegaShared {
    MegaOp op[2];        // double-buffered: the next descriptor is fetched while the current op runs
    float red[32];
    unsigned long long attn_bar;
    int s_last;
    float cv[kMW];
    int ci[kMW];
};

// All 32 lanes of the producer warp issue copies, each for the sequence numbers lane, lane + 32, ... of every GEMV: a single
// thread cannot re-issue a freed slot fast enough (16 consumer warps free one every ~50 ns).
__device__ __forceinline__ void producer_loop(const MegaCtl& c, uint8_t* smem, int n_act, int lane) {
    uint8_t* const ring = smem + c.region_bytes;
    const uint32_t ring_u = smem_u32(ring);
    const uint32_t full0 = smem_u32(ring + (size_t)c.nslots * c.slot_bytes), empty0 = full0 + c.nslots * 8;
    const uint64_t pol = policy_evict_first();
    const unsigned int nslots = (unsigned int)c.nslots;
    unsigned int seq_base = 0;
    for (int s = 0; s < n_act; s++) {
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(c.streams + s));
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(c.streams + s) + 1);
        const int total = (int)lo.z, per_cta = (int)lo.w, per_warp = (int)hi.x;
        const uint32_t bt = hi.y;
        const int i0 = blockIdx.x * per_cta, n_cta = min(total, i0 + per_cta) - i0;
        if (n_cta <= 0) continue;
        const int w_full = n_cta / per_warp, n_part = n_cta - w_full * per_warp;
        const int n_big = n_part * (w_full + 1);   // sequence numbers of the rounds that include the partial warp
        const uint8_t* src0 = reinterpret_cast<const uint8_t*>(((unsigned long long)lo.y << 32) | lo.x) + (size_t)i0 * bt;
        // Batches of 32 consecutive sequence numbers, the warp re-converging after each: lanes never drift two ring laps
        // apart (nslots >= 32), so a slot's empty-barrier parity always means the lap the lane is waiting for.
        for (int q0 = 0; q0 < n_cta; q0 += 32) {
            const int q = q0 + lane;
            if (q < n_cta) {
                int j, w;
                if (q < n_big) { j = q / (w_full + 1); w = q - j * (w_full + 1); }
                else { const int q2 = q - n_big; const int jj = q2 / w_full; j = n_part + jj; w = q2 - jj * w_full; }
                const unsigned int seq = seq_base + (unsigned int)q, lap = seq / nslots, slot = seq - lap * nslots;
                const uint32_t fb = full0 + slot * 8;
                mbar_wait_spin(empty0 + slot * 8, (lap & 1u) ^ 1u);   // a fresh mbarrier passes a wait on the phase before its first
                mbar_expect_tx(fb, bt);
                bulk_g2s_hint(ring_u + slot * c.slot_bytes, src0 + (size_t)(w * per_warp + j) * bt, bt, fb, pol);
            }
            __syncwarp();
        }
        seq_base += (unsigned int)n_cta;
    }
}

// ---- fused activation prologue of a GEMV, entirely in registers (gemv_mma_kernel's, as always-inlined helpers: a lambda that is
// not inlined would take the address of the register arrays and push them to local memory) ---------------------------------
template <int TYPE>
__device__ __forceinline__ int mega_f4(int b, int h, int lane) { return TYPE == kQ6_K ? 64 * b + lane + 32 * h : 64 * b + 2 * lane + h; }

// MODE 0: x = [rmsnorm_w2](a [+ r]);  1: Gemma-3 post-norm chain x = [rmsnorm_w2](rmsnorm_w1(a) + r);  2: x = silu(a) * a[K..].
// The x vector is streamed one 256-element block per warp at a time: load, residual add, gain, digit fragments -- nothing is
// held in a register array across the block-wide reductions and the f64 rsqrt (the old array form was parked in local memory
// by ptxas, and local memory is an L2 round trip when most of the SM's L1 is carved out as shared memory).
// The final RMSNorm scale is a scalar: the fragments are built from u = v * w2 and s = 1/rms(v) goes into the per-block
// inverse scale (deq(W).(s u) = s deq(W).u, and every term of a block-tile's sum carries xinv[b]).
template <int TYPE>
__device__ __forceinline__ F8 mega_ld8(const float4* p4, int b, int lane, int K4) {   // constants (norm gains): read-only path
    F8 x;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int i = mega_f4<TYPE>(b, h, lane);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < K4) v = __ldg(p4 + i);
        x.v[4 * h] = v.x; x.v[4 * h + 1] = v.y; x.v[4 * h + 2] = v.z; x.v[4 * h + 3] = v.w;
    }
    return x;
}

// The lane's 8 elements of block b of up to three flagged sources at once (two planes of the activation and the residual,
// each with its own epoch; a null source reads as zeros): all 16-byte loads are issued together and re-issued until every
// pair carries its epoch -- one L2 round trip once the data is there, however many sources.
struct LLSrc { const uint2* p; uint32_t epoch; };

template <int TYPE, int NS>
__device__ __forceinline__ void mega_ld8_group(const LLSrc (&src)[NS], int b, int lane, int K4, F8 (&out)[NS]) {
    const int i0 = mega_f4<TYPE>(b, 0, lane), i1 = mega_f4<TYPE>(b, 1, lane);
    const bool in0 = i0 < K4, in1 = i1 < K4;
    uint32_t d[NS][8], f[NS][8], spins = 0;
#pragma unroll
    for (int k = 0; k < NS; k++) {
#pragma unroll
        for (int e = 0; e < 8; e++) { d[k][e] = 0u; f[k][e] = src[k].epoch; }
    }
    bool ok;
    do {
#pragma unroll
        for (int k = 0; k < NS; k++) {
            if (src[k].p) {
                const uint2* base = src[k].p;
                if (in0) {
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d[k][0]), "=r"(f[k][0]), "=r"(d[k][1]), "=r"(f[k][1]) : "l"(base + 4 * i0) : "memory");
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d[k][2]), "=r"(f[k][2]), "=r"(d[k][3]), "=r"(f[k][3]) : "l"(base + 4 * i0 + 2) : "memory");
                }
                if (in1) {
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d[k][4]), "=r"(f[k][4]), "=r"(d[k][5]), "=r"(f[k][5]) : "l"(base + 4 * i1) : "memory");
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d[k][6]), "=r"(f[k][6]), "=r"(d[k][7]), "=r"(f[k][7]) : "l"(base + 4 * i1 + 2) : "memory");
                }
            }
        }
        ok = true;
#pragma unroll
        for (int k = 0; k < NS; k++) {
#pragma unroll
            for (int e = 0; e < 8; e++) ok = ok && f[k][e] == src[k].epoch;
        }
        if (++spins > (1u << 24)) __trap();   // a lost producer traps instead of hanging the GPU
    } while (!ok);
#pragma unroll
    for (int k = 0; k < NS; k++) {
#pragma unroll
        for (int e = 0; e < 8; e++) out[k].v[e] = __uint_as_float(d[k][e]);
    }
}

// activation (planes added in plane order) and residual of block b: x = a, y = r
template <int TYPE>
__device__ __forceinline__ void mega_ld_ar(const MegaVec& a, int a_off, uint32_t ea, const MegaVec& r, bool has_r, uint32_t er, int b, int lane, int K4,
                                           F8& x, F8& y) {
    const LLSrc src[3] = {{a.p + a_off, ea}, {a.planes > 1 ? a.p + (size_t)a.stride + a_off : nullptr, ea}, {has_r ? r.p : nullptr, er}};
    F8 o[3];
    mega_ld8_group<TYPE, 3>(src, b, lane, K4, o);
    x = o[0];
    if (a.planes > 1) {
#pragma unroll
        for (int e = 0; e < 8; e++) x.v[e] += o[1].v[e];
    }
    for (int pl = 2; pl < a.planes; pl++) {   // row tiles spread over more than two CTAs: small matrices only
        const LLSrc s1[1] = {{a.p + (size_t)pl * a.stride + a_off, ea}};
        F8 t[1];
        mega_ld8_group<TYPE, 1>(s1, b, lane, K4, t);
#pragma unroll
        for (int e = 0; e < 8; e++) x.v[e] += t[0].v[e];
    }
    y = o[2];
}

// residual add (y), running sum of squares, gain (w), digit fragments of one 256-element block held in registers
template <int TYPE, int MODE>
__device__ __forceinline__ float mega_finish_block(F8 x, const F8 y, const F8 w, int b, bool has_r, bool has_g, uint2* so, float* so_plain,
                                                   uint32_t so_epoch, int K4, float ss, uint4* xf, uint32_t* xm, float* xinv, int lane) {
    if (MODE == 2) {   // silu_generic.go:22-31: sigmoid in f64, float32(g*sig)*u
#pragma unroll
        for (int e = 0; e < 8; e++) x.v[e] = silu_mul(x.v[e], y.v[e]);
    } else {
        if (has_r) {
#pragma unroll
            for (int e = 0; e < 8; e++) x.v[e] += y.v[e];
        }
        if (so || so_plain) {   // CTA 0 only: the residual stream for the ops (and the host) that read it later
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int i = mega_f4<TYPE>(b, h, lane);
                if (i < K4) {
                    if (so) {
                        st_pair2(so + 4 * i, x.v[4 * h], x.v[4 * h + 1], so_epoch);
                        st_pair2(so + 4 * i + 2, x.v[4 * h + 2], x.v[4 * h + 3], so_epoch);
                    }
                    if (so_plain) reinterpret_cast<float4*>(so_plain)[i] = make_float4(x.v[4 * h], x.v[4 * h + 1], x.v[4 * h + 2], x.v[4 * h + 3]);
                }
            }
        }
        if (has_g) {
#pragma unroll
            for (int e = 0; e < 8; e++) { ss = fmaf(x.v[e], x.v[e], ss); x.v[e] *= w.v[e]; }
        }
    }
    if (TYPE == kQ4_K || TYPE == kQ5_K) frags_q4k_i8(x, b, lane, xf, xm, xinv);
    else if (TYPE == kQ6_K) frags_q6k_i8(x, b, lane, xf, xinv);
    else frags_q40_i8(x, b, lane, 64 * b + 2 * lane < K4, reinterpret_cast<uint2*>(xf), reinterpret_cast<int*>(xm), xinv);
    return ss;
}

template <int TYPE, int MODE>
__device__ __forceinline__ void mega_build_frags(const MegaGemv& g, uint32_t epoch_base, uint4* xf, uint32_t* xm, float* xinv, float* red, int warp,
                                                 int lane) {
    const int K = g.K;
    const int nxb = (K + 255) >> 8;
    const int K4 = K >> 2;
    const uint32_t ea = epoch_base + (uint32_t)g.a.tag_op, er = epoch_base + (uint32_t)g.r.tag_op, eo = epoch_base + (uint32_t)g.tag_op;
    const bool has_r = MODE != 2 && g.r.p != nullptr;
    const float4* g1 = reinterpret_cast<const float4*>(g.w1);
    const float4* g4 = reinterpret_cast<const float4*>(g.w2);
    uint2* so = (MODE != 2 && blockIdx.x == 0) ? g.sum_out : nullptr;
    float* so_plain = (MODE != 2 && blockIdx.x == 0) ? g.sum_plain : nullptr;
    float s1 = 1.0f;
    const MegaVec none{};
    if (MODE == 1) {   // first norm of the chain: needs the whole vector's sum of squares before anything else
        float q = 0.0f;
        for (int b = warp; b < nxb; b += kMW) {
            F8 x, y;
            mega_ld_ar<TYPE>(g.a, 0, ea, none, false, 0u, b, lane, K4, x, y);
#pragma unroll
            for (int e = 0; e < 8; e++) q = fmaf(x.v[e], x.v[e], q);
        }
        s1 = inv_rms(csum(q, red), K, g.eps);
    }
    float ss = 0.0f;
    for (int b = warp; b < nxb; b += kMW) {
        F8 x, y, w = {};
        if (MODE == 2) {
            F8 dummy;
            mega_ld_ar<TYPE>(g.a, 0, ea, none, false, 0u, b, lane, K4, x, dummy);
            mega_ld_ar<TYPE>(g.a, K, ea, none, false, 0u, b, lane, K4, y, dummy);
        } else {
            mega_ld_ar<TYPE>(g.a, 0, ea, g.r, has_r, er, b, lane, K4, x, y);
            if (MODE == 1) {
                const F8 g1v = mega_ld8<TYPE>(g1, b, lane, K4);
#pragma unroll
                for (int e = 0; e < 8; e++) x.v[e] = x.v[e] * s1 * g1v.v[e];
            }
            if (g4) w = mega_ld8<TYPE>(g4, b, lane, K4);
        }
        ss = mega_finish_block<TYPE, MODE>(x, y, w, b, has_r, g4 != nullptr, so, so_plain, eo, K4, ss, xf, xm, xinv, lane);
    }
    if (MODE != 2 && g4) {
        const float s2 = inv_rms(csum(ss, red), K, g.eps);   // csum's barriers also order the xinv writes above
        if (lane == 0)   // this lane wrote the inverse scales of the warp's blocks
            for (int b = warp; b < nxb; b += kMW) xinv[b] *= s2;
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

// ---- one GEMV op: fused prologue in registers, the warp's run of block-tiles from the ring, partial-sum exchange, epilogue -----
// Returns the number of block-tiles this CTA consumed (the caller's running sequence number advances by it).
template <int TYPE>
__device__ __noinline__ int gemv_phase(const MegaCtl& c, MegaShared* sh, const MegaOp* op, uint8_t* smem, unsigned int seq_base, int oi, int warp,
                                       int lane, uint32_t epoch_base) {
    // The op descriptor lives in shared memory: every stage below re-reads what it needs after the barrier that precedes it,
    // so that nothing but the stage's own working set is live in registers (the Q6_K tile alone takes ~120).
    const MegaGemv& g = op->g;
    float* red = sh->red;
    constexpr int BT = bt_bytes(TYPE);
    const int i0 = blockIdx.x * g.per_cta, i1 = min(g.total, i0 + g.per_cta);
    const int n_cta = i1 - i0;
    if (n_cta <= 0) {   // CTA-uniform: no block-tiles of this matrix here
        if (g.head && threadIdx.x == 0) {
            c.cand_v[blockIdx.x] = -FLT_MAX;
            c.cand_i[blockIdx.x] = 0x7fffffff;
        }
        return 0;
    }

    // ---- fused prologue, in registers: warp w builds super-blocks w, w+16, ... of x (zb_stream.cuh Prologue semantics)
    {
        uint4* xf = reinterpret_cast<uint4*>(smem + g.xf_off);
        uint32_t* xm = reinterpret_cast<uint32_t*>(smem + g.xm_off);
        float* xinv = reinterpret_cast<float*>(smem + g.xinv_off);
        if (g.swiglu) mega_build_frags<TYPE, 2>(g, epoch_base, xf, xm, xinv, red, warp, lane);
        else if (g.w1) mega_build_frags<TYPE, 1>(g, epoch_base, xf, xm, xinv, red, warp, lane);
        else mega_build_frags<TYPE, 0>(g, epoch_base, xf, xm, xinv, red, warp, lane);
    }
    csync();
    mega_stamp(c.trace, oi, 1);

    // ---- main loop: this warp's run of block-tiles, one ring slot each (gemv_mma_kernel's arithmetic and order)
    const int nb = g.nb;
    float* part = reinterpret_cast<float*>(smem + g.part_off);
    const int tau_first = i0 / nb;
    const int t = lane & 3;
    {
        const int w_full = n_cta / g.per_warp, n_part = n_cta - w_full * g.per_warp;
        const int n_my = warp < w_full ? g.per_warp : (warp == w_full ? n_part : 0);
        const int r0 = i0 + warp * g.per_warp;
        uint8_t* const ring = smem + c.region_bytes;
        const uint32_t full0 = smem_u32(ring + (size_t)c.nslots * c.slot_bytes), empty0 = full0 + c.nslots * 8;
        const uint4* xf = reinterpret_cast<const uint4*>(smem + g.xf_off);
        const uint32_t* xm = reinterpret_cast<const uint32_t*>(smem + g.xm_off);
        const float* xinv = reinterpret_cast<const float*>(smem + g.xinv_off);
        const float wlo = t == 0 ? 16777216.0f : (t == 1 ? 256.0f : 0.0f), whi = t == 0 ? 65536.0f : (t == 1 ? 1.0f : 0.0f);
        const uint32_t q6selA = 24u - 8u * (uint32_t)(t >> 1), q6selB = 8u - 8u * (uint32_t)(t >> 1);
        float tot[4] = {0.f, 0.f, 0.f, 0.f};
        int tau = r0 / nb, b = r0 - tau * nb;
        int cur_tau = -1;
        for (int j = 0; j < n_my; j++) {
            const unsigned int seq = seq_base + (unsigned int)(j * w_full + min(j, n_part) + warp);
            const unsigned int lap = seq / (unsigned int)c.nslots, slot = seq - lap * (unsigned int)c.nslots;
            mbar_wait_spin(full0 + slot * 8, lap & 1u);
            if (tau != cur_tau) {
                if (cur_tau >= 0) mega_flush(tot, part, cur_tau - tau_first, g.slots, warp - (max(i0, cur_tau * nb) - i0) / g.per_warp, lane);
                cur_tau = tau;
                tot[0] = tot[1] = tot[2] = tot[3] = 0.0f;
            }
            const uint8_t* bt = ring + (size_t)slot * c.slot_bytes;
            if (TYPE == kQ4_K || TYPE == kQ5_K)
                block_tile_q4k_i8<TYPE == kQ5_K>(bt, xf + (size_t)b * 96, xm + b * kXmWords, xinv[b], tot, lane, wlo, whi);
            else if (TYPE == kQ6_K)
                block_tile_q6k_i8(bt, xf + (size_t)b * 96, xinv[b], tot, lane, q6selA, q6selB, (t & 1) ? 256.0f : 16777216.0f,
                                  (t & 1) ? 1.0f : 65536.0f);
            else
                block_tile_q40_i8(bt, reinterpret_cast<const uint2*>(xf) + (size_t)b * 64, reinterpret_cast<const int*>(xm) + b * 16,
                                  xinv[b >> 1], tot, lane, wlo, whi);
            if (++b == nb) { b = 0; tau++; }
            __syncwarp();   // every lane has read the slot
            if (lane == 0) mbar_arrive(empty0 + slot * 8);
        }
        if (cur_tau >= 0) mega_flush(tot, part, cur_tau - tau_first, g.slots, warp - (max(i0, cur_tau * nb) - i0) / g.per_warp, lane);
        (void)BT;
    }
    mega_stamp(c.trace, oi, 2);
    csync();
    mega_stamp(c.trace, oi, 3);

    // ---- per row tile: sum the warps' partials in slot order and publish.  A tile that straddles CTAs is published in
    // planes: the CTA holding its k-th part writes plane k, the first one also zero-fills the planes nobody else writes;
    // the readers add the planes in order.  Tiles of ops with an epilogue (SwiGLU pairs, lm_head) are never split.
    const int M = g.M;
    const uint32_t eo = epoch_base + (uint32_t)g.tag_op;
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
        __syncwarp();
        const float up = __shfl_down_sync(0xffffffffu, v, 1);
        if (valid) {
            const int grow = tt * 16 + row;
            if (g.pairs) {
                if (!(row & 1) && grow + 1 < M) st_pair(g.y + (grow >> 1), silu_mul(v, up), eo);
            } else if (grow < M) {
                if (g.head) {
                    if (cap > 0.0f) v = softcap_apply(v, cap, inv_cap);
                    g.y_plain[grow] = v;
                    if (v > best_v || (v == best_v && grow < best_i)) { best_v = v; best_i = grow; }
                } else {
                    const int c_first = (tt * nb) / g.per_cta, c_last = ((tt + 1) * nb - 1) / g.per_cta;
                    const int mypart = blockIdx.x - c_first;
                    st_pair(g.y + (size_t)mypart * g.y_stride + grow, v, eo);
                    if (mypart == 0)
                        for (int pl = c_last - c_first + 1; pl < g.y_planes; pl++) st_pair(g.y + (size_t)pl * g.y_stride + grow, 0.0f, eo);
                }
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
        csync();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kMW; w++)
                if (sh->cv[w] > best_v || (sh->cv[w] == best_v && sh->ci[w] < best_i)) { best_v = sh->cv[w]; best_i = sh->ci[w]; }
            c.cand_v[blockIdx.x] = best_v;
            c.cand_i[blockIdx.x] = best_i;
        }
    }
    return n_cta;
}

__device__ __noinline__ void attn_phase(MegaShared* sh, const MegaOp* op, uint8_t* smem, uint32_t& attn_parity, uint32_t epoch_base, int oi) {
    const AttnArgs& a = op->a;
    const int warp = threadIdx.x >> 5;
    const int G = gridDim.x;
    const uint32_t attn_bar = smem_u32(&sh->attn_bar);
    const int pos = *a.pos_ptr;
    const int len = pos + 1, nsplits = (len + a.chunk - 1) / a.chunk, items = a.nkv * nsplits;
    const int rep = a.nq / a.nkv;
    const uint32_t ei = epoch_base + (uint32_t)a.qkv_tag_op, eo = epoch_base + (uint32_t)oi;
    uint32_t par = attn_parity;
    for (int it = blockIdx.x; it < items; it += G) {
        const int kvh = it / nsplits, split = it - kvh * nsplits;
        if (warp < kMegaAttnWarps) {
            switch (a.hd) {
                case 32: attn_item_rep<1>(rep, &a, kvh, split, pos, smem, attn_bar, par, &sh->s_last, ei, eo); break;
                case 64: attn_item_rep<2>(rep, &a, kvh, split, pos, smem, attn_bar, par, &sh->s_last, ei, eo); break;
                case 128: attn_item_rep<4>(rep, &a, kvh, split, pos, smem, attn_bar, par, &sh->s_last, ei, eo); break;
                case 256: attn_item_rep<8>(rep, &a, kvh, split, pos, smem, attn_bar, par, &sh->s_last, ei, eo); break;
            }
        }
        par ^= 1u;
    }
    attn_parity = par;
}

__device__ __noinline__ void embed_phase(const MegaOp* op, uint32_t epoch) {
    // token select + embedding row gather, bit-exact dequantisation (+ Gemma scale): arch_llama.go:246-342, arch_gemma.go:38
    const MegaEmbed& e = op->e;
    const int fi = *e.feed_idx;
    int tok = fi < *e.feed_len ? e.feed[fi] : *e.last;
    if (tok < 0) tok = 0;
    if (tok >= e.vocab) tok = e.vocab - 1;
    const int64_t base = (int64_t)tok * e.hidden;
    for (int i = blockIdx.x * kMT + threadIdx.x; i < e.hidden; i += gridDim.x * kMT) {
        const float v = deq_raw(e.type, e.table, base + i);
        st_pair(e.out + i, e.scale > 0.0f ? v * e.scale : v, epoch);
    }
}

__device__ __noinline__ void final_phase(const MegaCtl* cp, const MegaOp* op, int with_head) {
    // argmax over the CTAs' candidates + step bookkeeping (sampling_helpers.go:11-45, tensor_cache.go:205-262 counters)
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const MegaCtl& c = *cp;
    const MegaFinal& f = op->f;
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
        *f.epoch_step += 1u;
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

__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel(const __grid_constant__ MegaCtl c, int with_head) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(16) MegaShared sh;

    const int G = gridDim.x;
    const int n_act = with_head ? c.n_streams : c.n_streams_nohead;
    constexpr int kOpWords = (int)(sizeof(MegaOp) / 4);
    if (threadIdx.x == 0) {
        const uint32_t full0 = smem_u32(smem + c.region_bytes + (size_t)c.nslots * c.slot_bytes);
        for (int s = 0; s < 2 * c.nslots; s++) mbar_init(full0 + s * 8, 1);   // full[nslots] then empty[nslots]
        mbar_init(smem_u32(&sh.attn_bar), 1);
        fence_barrier_init();
    }
    if (threadIdx.x < kOpWords) reinterpret_cast<uint32_t*>(&sh.op[0])[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t*>(c.ops) + threadIdx.x);
    __syncthreads();   // the only CTA-wide barrier: from here on the producer warp and the consumers go their own ways

    if (threadIdx.x >= kMT) {   // producer warp: the weights are constants, stream them as far ahead as the ring allows
        producer_loop(c, smem, n_act, threadIdx.x - kMT);
        return;
    }

    const unsigned int bar_base = (unsigned int)__ldcg(c.step) * (unsigned int)c.n_barriers * (unsigned int)G;
    const uint32_t epoch_base = 1u + __ldcg(c.epoch_step) * (uint32_t)c.n_ops;
    unsigned int bar_k = 0, seq_base = 0;
    uint32_t attn_parity = 0;

    for (int oi = 0; oi < c.n_ops; oi++) {
        const MegaOp* op = &sh.op[oi & 1];
        // next descriptor: the load is issued now, its value parked in a register and stored after the op's work
        uint32_t nxt = 0u;
        if (oi + 1 < c.n_ops && threadIdx.x < kOpWords) nxt = __ldg(reinterpret_cast<const uint32_t*>(c.ops + oi + 1) + threadIdx.x);
        const int kind = op->kind;
        mega_stamp(c.trace, oi, 0);
        // Launder the thread coordinates and the shared-memory base once per op: without this the compiler hoists every
        // lane- / warp-dependent address and constant of all GEMV variants out of the op loop and keeps them alive across it.
        int warp_v = threadIdx.x >> 5, lane_v = threadIdx.x & 31;
        uint8_t* smem_v = smem;
        asm volatile("" : "+r"(warp_v), "+r"(lane_v), "+l"(smem_v));
        if (kind == kMegaGemv) {
            if (!(op->g.head && !with_head)) {
                int n = 0;
                switch (op->g.type) {
                    case kQ4_K: n = gemv_phase<kQ4_K>(c, &sh, op, smem_v, seq_base, oi, warp_v, lane_v, epoch_base); break;
                    case kQ5_K: n = gemv_phase<kQ5_K>(c, &sh, op, smem_v, seq_base, oi, warp_v, lane_v, epoch_base); break;
                    case kQ6_K: n = gemv_phase<kQ6_K>(c, &sh, op, smem_v, seq_base, oi, warp_v, lane_v, epoch_base); break;
                    default: n = gemv_phase<kQ4_0>(c, &sh, op, smem_v, seq_base, oi, warp_v, lane_v, epoch_base); break;
                }
                seq_base += (unsigned int)n;
            }
        } else if (kind == kMegaAttn) {
            attn_phase(&sh, op, smem, attn_parity, epoch_base, oi);
        } else if (kind == kMegaEmbed) {
            embed_phase(op, epoch_base + (uint32_t)oi);
        } else {
            final_phase(&c, op, with_head);
        }
        mega_stamp(c.trace, oi, 4);
        const int barrier = op->barrier;
        if (oi + 1 < c.n_ops && threadIdx.x < kOpWords) reinterpret_cast<uint32_t*>(&sh.op[(oi + 1) & 1])[threadIdx.x] = nxt;
        if (barrier) {
            bar_k++;
            grid_barrier(c.bar_counter, bar_base + bar_k * (unsigned int)G);
        } else {
            csync();   // everybody is done with this op's scratch region; the next descriptor is in place
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
    // 196 KB of the SM's 256 KB as shared memory: the rest stays L1 for the op table, norm gains and the few local spills
    e = cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 85);
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
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_mega_kernel, kMegaThreads, kMegaSmem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
    *out_ctas = sms < ZB_SMS ? sms : ZB_SMS;
    return 0;
}

int mega_launch(const MegaCtl& ctl, int ctas, int with_head, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas, 1, 1);
    cfg.blockDim = dim3(kMegaThreads, 1, 1);
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
