// TMA-streamed fused dequant-GEMV for batch-1 decode (sm_100a).
//
//   y[M] = deq(W[M,K]) . x[K],   x = prologue(a, r, w1, w2)   (zb_stream.cuh)
//
// Why this shape.  At B200 rates (6.5 TB/s over 148 SMs = 23 B/clk/SM) a 4-bit
// GEMV is co-limited by HBM and by instruction issue (41 weights/clk/SM against
// 128 lane-ops/clk/SM), and every decode matrix of a 1B-8B model is smaller than
// the bandwidth-latency product, so a kernel that only starts loading after it
// has been launched and has seen its input never reaches the roofline.  Hence:
//   * weights travel global -> shared through the TMA engine (cp.async.bulk +
//     mbarrier): each warp owns a private 2-4 stage ring of (R rows x K-slab)
//     tiles, ~100 KB in flight per SM independent of register pressure;
//   * the first ring fill is issued BEFORE griddepcontrol.wait: with programmatic
//     dependent launch the next GEMV's weights stream in while the previous
//     kernel is still computing (weights are constants, only x is a dependency);
//   * x is built once per CTA in shared memory by a fused prologue (RMSNorm /
//     Add+RMSNorm / Norm+Add+Norm / SwiGLU / MoE combine), stored in the order
//     the format consumes it and XOR-swizzled so the 128-bit reads are
//     conflict-free; one x chunk (32 floats in registers) serves R rows;
//   * dequantisation is in registers: nibbles are masked four at a time, PRMT
//     drops each into the mantissa of 128.0f (exact small integers), and the
//     bias subtract + multiply-accumulate run as packed FADD2/FFMA2;
//   * rows meet in a warp-shuffle reduction over the lanes that share them.
// Determinism: the lane/chunk mapping is fixed by (type, M, K), so results are
// bit-reproducible run to run.
//
// Reference semantics replaced: Engine.MatMul on quantized storage + the fused
// providers around it: gemm_q4.cu:48-120,
// gemv_q4k.cu:68-142, gemv_q5k.cu:68-156, gemv_q6k.cu:45-131, gemm_q8.cu:24-118,
// fused_add_rmsnorm.cu:17-55, fused_norm_add.cu:11-49, rmsnorm.cu:11-61,
// fused_swiglu.cu:11-22, layers/core/moe.go:463-485.
#include <stdlib.h>

#include "zb_prologue.cuh"
#include "zb200.h"

namespace {

using namespace zb;

constexpr int kSWarps = 8;
constexpr int kSThreads = kSWarps * 32;
constexpr int kCtasPerSm = 3;             // 24 warps / SM: the dequant is issue-bound, it needs the thread-level parallelism
constexpr int kMaxStages = 4;
constexpr int kSmemTotal = 225 * 1024;    // usable shared memory per SM, split between the co-resident CTAs

// unit_w / xpos / build_x: zb_prologue.cuh
__host__ __device__ inline int unit_main_bytes(int t, int units) { return stream_main_bytes(t, units * (unit_w(t) / 32)); }
__host__ __device__ inline int unit_aux_bytes(int t, int units) { return stream_aux_bytes(t, units * (unit_w(t) / 32)); }

struct SGeom {
    int U;                   // units per row
    int lpr;                 // lanes sharing a row (4..32)
    int rows_pass;           // rows per tile = (32/lpr) * R
    int upl;                 // units per lane per slab
    int slab_units, n_slabs;
    int contig;              // slab == whole row: a tile is one contiguous span of the matrix
    int row_main, row_aux;   // bytes per matrix row
    int slab_main;           // main bytes of one full row-slab
    int slab_main_cap, slab_aux_cap;  // per-row slot sizes inside a stage (slab mode)
    int stage_main, stage_bytes;      // aux region starts at stage_main
    int stages, n_tiles;
    int xsum_off, ring_off, bar_off, smem_bytes, ctas_per_sm;
    int mma, xfrag_off;      // tensor-core (mma.sync) dequant path: 16 rows per warp tile, x pre-split into fp16 B fragments
};

struct Indirect {            // MoE: blockIdx.y = slot k, expert = sel[k]
    const int* sel;
    long long main_stride, aux_stride;
    int a_stride, y_stride;
    int swiglu_pairs;        // epilogue: rows (2i, 2i+1) hold (gate_i, up_i); store y[i] = silu(gate_i) * up_i
    // fused tensor-parallel exchange (producer side): rows go to this rank's slot on every peer, last CTA publishes the epoch
    int n_peers, site, sites_per_step;
    float* peer_out[8];
    unsigned int* peer_flag[8];
    const int* epoch_base;
    int* ticket;
    // consumer side: flags to wait for before the prologue reads the slots
    const unsigned int* wait_flags;
    const int* wait_epoch_base;
    int n_wait, wait_site, wait_sites_per_step;
};


// ---- in-register dequantisation ------------------------------------------------
// four masked bytes (each < 128) -> two f32x2 pairs holding 128+b, then minus `bias`
#define ZB_C43 0x43000000u
__device__ __forceinline__ void bytes_to_pairs(uint32_t m, uint64_t nbias, uint64_t& p0, uint64_t& p1) {
    p0 = add2(pack2u(__byte_perm(m, ZB_C43, 0x7044), __byte_perm(m, ZB_C43, 0x7144)), nbias);
    p1 = add2(pack2u(__byte_perm(m, ZB_C43, 0x7244), __byte_perm(m, ZB_C43, 0x7344)), nbias);
}
// acc += dot(4 masked bytes, 4 x values held as two pairs)
__device__ __forceinline__ uint64_t dot4(uint32_t m, uint64_t nbias, uint64_t x0, uint64_t x1, uint64_t acc) {
    uint64_t p0, p1;
    bytes_to_pairs(m, nbias, p0, p1);
    acc = fma2(p0, x0, acc);
    return fma2(p1, x1, acc);
}
// (Folding the float trick's 128 into one 128*sum(x) correction per scale group would save a FADD2 per two weights, but it
// was measured to cost precision -- |err| up to 2.8e-5, outside the reference's 1e-5 + 1e-4|ref| bound -- for ~4 % speed.)
// four 16-B groups (a quarter of a 64-float unit / half of a 32-float unit) of x as 8 f32x2 pairs
__device__ __forceinline__ void ld_xq(uint64_t (&xv)[8], const ulonglong2* xu, int g0, int sw) {
    (void)sw;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        ulonglong2 t = xu[g0 + j];
        xv[2 * j] = t.x;
        xv[2 * j + 1] = t.y;
    }
}

// All R rows of one unit `ul` of the slab: x is read from shared memory once and serves every row.
// The loops over the 16-weight quarters of a unit are deliberately NOT unrolled: the dequant is issue-bound and
// a fully unrolled body (600+ instructions) overflows the per-SMSP L0 instruction cache -- ncu showed
// stall_no_instruction as the top stall -- so the body is kept near 250 instructions.
template <int TYPE, int R>
__device__ __forceinline__ void unit_dot(const uint8_t* const (&rowm)[R], const uint8_t* const (&rowa)[R], int ul, const float* xs_u, int sw,
                                         float4 xsm, float (&acc)[R], int lane) {
    const ulonglong2* xu = reinterpret_cast<const ulonglong2*>(xs_u);
    uint64_t xv[8];
    if (TYPE == kQ4_0) {
        const uint64_t nb = pack2(-136.0f, -136.0f);  // 128 (float trick) + 8 (Q4_0 offset)
        uint64_t xw[8];
        ld_xq(xv, xu, 0, sw);   // low nibbles <-> x[0..15]
        ld_xq(xw, xu, 4, sw);   // high nibbles <-> x[16..31]
#pragma unroll
        for (int j = 0; j < R; j++) {
            uint4 q = *reinterpret_cast<const uint4*>(rowm[j] + ul * 16);
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint64_t a = 0ull;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                a = dot4(w[i] & 0x0F0F0F0Fu, nb, xv[2 * i], xv[2 * i + 1], a);
                a = dot4((w[i] >> 4) & 0x0F0F0F0Fu, nb, xw[2 * i], xw[2 * i + 1], a);
            }
            acc[j] += sum2(a) * h2f(*reinterpret_cast<const uint16_t*>(rowa[j] + ul * 2));
        }
    } else if (TYPE == kQ8_0) {
        const uint64_t nb = pack2(-8388736.0f, -8388736.0f);  // 2^23 + 128
        const int hs = (lane >> 2) & 1;                        // stagger the two 16-B halves: conflict-free at 32-B lane stride
        uint64_t a[R];
#pragma unroll
        for (int j = 0; j < R; j++) a[j] = 0ull;
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
            const int hh = t ^ hs;
            ld_xq(xv, xu, 4 * hh, sw);
#pragma unroll
            for (int j = 0; j < R; j++) {
                uint4 q = *reinterpret_cast<const uint4*>(rowm[j] + ul * 32 + hh * 16);
                uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t u = w[i] ^ 0x80808080u;  // int8 -> biased uint8
                    uint64_t p0 = add2(pack2u(__byte_perm(u, 0x4B000000u, 0x7440), __byte_perm(u, 0x4B000000u, 0x7441)), nb);
                    uint64_t p1 = add2(pack2u(__byte_perm(u, 0x4B000000u, 0x7442), __byte_perm(u, 0x4B000000u, 0x7443)), nb);
                    a[j] = fma2(p0, xv[2 * i], a[j]);
                    a[j] = fma2(p1, xv[2 * i + 1], a[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; j++) acc[j] += sum2(a[j]) * h2f(*reinterpret_cast<const uint16_t*>(rowa[j] + ul * 2));
    } else if (TYPE == kQ4_K || TYPE == kQ5_K) {
        constexpr int BB = TYPE == kQ4_K ? 144 : 176;
        const int boff = (ul >> 2) * BB, g = ul & 3;
        const uint64_t nb = pack2(-128.0f, -128.0f);
        // 6-bit (scale, min) of sub-blocks 2g and 2g+1 decoded two at a time from the packed 12 bytes (gemv_q4k.cu:38-56)
        const int sh = (g & 1) * 16;
        float dsA[R], dsB[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            uint4 hdr = *reinterpret_cast<const uint4*>(rowm[j] + boff);
            float d = h2f((uint16_t)(hdr.x & 0xFFFFu)), dmin = h2f((uint16_t)(hdr.x >> 16));
            uint32_t h0 = (hdr.y >> sh) & 0xFFFFu, h1 = (hdr.z >> sh) & 0xFFFFu, h2 = (hdr.w >> sh) & 0xFFFFu;
            uint32_t sc2 = g < 2 ? (h0 & 0x3F3Fu) : ((h2 & 0x0F0Fu) | ((h0 >> 2) & 0x3030u));
            uint32_t mn2 = g < 2 ? (h1 & 0x3F3Fu) : (((h2 >> 4) & 0x0F0Fu) | ((h1 >> 2) & 0x3030u));
            dsA[j] = d * (float)(sc2 & 0xFFu);  // exact products (fp16 x 6-bit)
            dsB[j] = d * (float)(sc2 >> 8);
            // sum (d*sc*q - dmin*m) x = d*sc*sum(q x) - dmin*m*sum(x): the min terms go in first
            acc[j] -= (dmin * (float)(mn2 & 0xFFu)) * xsm.x + (dmin * (float)(mn2 >> 8)) * xsm.y;
        }
        const uint8_t* qp[R];
#pragma unroll
        for (int j = 0; j < R; j++) qp[j] = rowm[j] + boff + 16 + g * 32;
        uint64_t xw[8];
#pragma unroll 1
        for (int hh = 0; hh < 2; hh++) {               // the two 16-byte halves of the group; low and high nibbles unrolled
            ld_xq(xv, xu + 4 * hh, 0, sw);              // x of the low-nibble weights 16hh .. 16hh+15 (sub-block 2g)
            ld_xq(xw, xu + 4 * hh, 8, sw);              // x of the high-nibble weights (sub-block 2g+1)
#pragma unroll
            for (int j = 0; j < R; j++) {
                uint4 q = *reinterpret_cast<const uint4*>(qp[j] + hh * 16);
                uint32_t w[4] = {q.x, q.y, q.z, q.w};
                uint32_t hl[4] = {0u, 0u, 0u, 0u}, hu[4] = {0u, 0u, 0u, 0u};
                if (TYPE == kQ5_K) {
                    uint4 h = *reinterpret_cast<const uint4*>(rowm[j] + boff + 144 + hh * 16);
                    uint32_t hv[4] = {h.x >> (2 * g), h.y >> (2 * g), h.z >> (2 * g), h.w >> (2 * g)};
#pragma unroll
                    for (int i = 0; i < 4; i++) { hl[i] = (hv[i] & 0x01010101u) << 4; hu[i] = (hv[i] & 0x02020202u) << 3; }
                }
                uint64_t a = 0ull, b = 0ull;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    a = dot4((w[i] & 0x0F0F0F0Fu) | hl[i], nb, xv[2 * i], xv[2 * i + 1], a);
                    b = dot4(((w[i] >> 4) & 0x0F0F0F0Fu) | hu[i], nb, xw[2 * i], xw[2 * i + 1], b);
                }
                acc[j] += dsA[j] * sum2(a) + dsB[j] * sum2(b);
            }
        }
    } else {  // kQ6_K, split layout: ql[128] qh[64] sc[16] per block, fp16 d in aux; unit = (half, lh)
        const int boff = (ul >> 2) * 208, sub = ul & 3, half = sub >> 1, lh = sub & 1;
        const uint64_t nb = pack2(-160.0f, -160.0f);  // 128 + 32
        uint64_t xw[8];
#pragma unroll 1
        for (int hp = 0; hp < 2; hp++) {               // hp = 0: q1 | q2 (low nibbles of ql), hp = 1: q3 | q4 (high nibbles)
            ld_xq(xv, xu + 8 * hp, 0, sw);
            ld_xq(xw, xu + 8 * hp, 4, sw);
#pragma unroll
            for (int j = 0; j < R; j++) {
                const uint8_t* blk = rowm[j] + boff;
                uint4 A = *reinterpret_cast<const uint4*>(blk + half * 64 + lh * 16);
                uint4 B = *reinterpret_cast<const uint4*>(blk + half * 64 + 32 + lh * 16);
                uint4 H = *reinterpret_cast<const uint4*>(blk + 128 + half * 32 + lh * 16);
                uint32_t a4[4] = {A.x, A.y, A.z, A.w}, b4[4] = {B.x, B.y, B.z, B.w}, h4[4] = {H.x, H.y, H.z, H.w};
                uint64_t ca = 0ull, cb = 0ull;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t hs = h4[i] >> (4 * hp);
                    uint32_t qa = ((a4[i] >> (4 * hp)) & 0x0F0F0F0Fu) | ((hs & 0x03030303u) << 4);
                    uint32_t qb = ((b4[i] >> (4 * hp)) & 0x0F0F0F0Fu) | ((hs & 0x0C0C0C0Cu) << 2);
                    ca = dot4(qa, nb, xv[2 * i], xv[2 * i + 1], ca);
                    cb = dot4(qb, nb, xw[2 * i], xw[2 * i + 1], cb);
                }
                const int8_t* sc = reinterpret_cast<const int8_t*>(blk + 192) + half * 8 + lh + 4 * hp;
                float d = h2f(*reinterpret_cast<const uint16_t*>(rowa[j] + (ul >> 2) * 2));
                acc[j] += (d * (float)sc[0]) * sum2(ca) + (d * (float)sc[2]) * sum2(cb);
            }
        }
    }
}

// ---- tensor-core path (mma.sync.m16n8k16, f16 x f16 -> f32) ---------------------------------------
// The CUDA-core dequant above is bound by instruction issue (ncu: 6-7 instructions per weight, issue slots 70 % busy at
// 3 TB/s).  Here the multiply-accumulate AND the int->float conversion move to the tensor cores: a nibble pair masked out
// of the quant word IS an fp16x2 operand (the subnormals n * 2^-24, exact), so a weight costs ~0.9 ALU instructions; the
// activations enter as three fp16 terms (x = h1 + h2 + h3, 33 mantissa bits) in three of the eight B columns, the f32
// accumulator keeps full precision, block scales are applied to the accumulator per 32-weight sub-block.
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t h2pack(float a, float b) {
    __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
// B fragments of x for every (unit, mma, split, t): b0 = (x[l0], x[l0+2]), b1 = (x[l0+1], x[l0+3]) -- the k order the nibble
// pairs of one quant word come out in.  xs holds f32 x in the unit-major padded layout of the CUDA-core path.
template <int TYPE>
__device__ void build_xfrag(const float* xs, uint2* xf, int K, float xscale) {
    constexpr int UW = unit_w(TYPE), NM = UW / 16;  // MMAs per unit
    const int U = K / UW;
    for (int i = threadIdx.x; i < U * NM * 4; i += kSThreads) {
        const int t = i & 3, m = (i >> 2) % NM, u = i / (4 * NM);
        // element offset (inside the unit's x) of the 4 weights lane t feeds to MMA m
        int l0;
        if (TYPE == kQ4_0) l0 = (m ? 16 : 0) + 4 * t;                  // word t: low nibbles x[0..15], high nibbles x[16..31]
        else l0 = (m >> 1) * 32 + 8 * t + 4 * (m & 1);                  // Q4_K/Q5_K: word 2t + (m&1); low-nibble half then high-nibble half
        const float* x = xs + u * (UW + 4) + l0;
        float v[4] = {x[0] * xscale, x[1] * xscale, x[2] * xscale, x[3] * xscale};
        uint2* dst = xf + ((size_t)(u * NM + m) * 3) * 4 + t;
#pragma unroll
        for (int sp = 0; sp < 3; sp++) {
            __half h[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                h[e] = __float2half_rn(v[e]);
                v[e] -= __half2float(h[e]);
            }
            uint2 o;
            o.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[2]) << 16);
            o.y = (uint32_t)__half_as_ushort(h[1]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
            dst[sp * 4] = o;
        }
    }
}

// One unit of 16 rows (rows g and g+8 per lane, g = lane>>2) on the tensor cores.  tot[0..1] = row g cols (2t, 2t+1),
// tot[2..3] = row g+8: columns 0,1,2 carry the three fp16 terms of x, the caller adds them up at the end of the row tile.
template <int TYPE>
__device__ __forceinline__ void unit_mma(const uint8_t* r0p, const uint8_t* r1p, const uint8_t* a0p, const uint8_t* a1p, int ul,
                                         const uint2* xf_u, float4 xsm, float (&tot)[4], int lane) {
    const int g = lane >> 2, t = lane & 3;
    (void)g;
    if (TYPE == kQ4_0) {
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(r0p + ul * 16 + 4 * t);
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(r1p + ul * 16 + 4 * t);
        uint2 b0 = make_uint2(0u, 0u), b1 = make_uint2(0u, 0u);
        if (lane < 12) { b0 = xf_u[(0 * 3 + (lane >> 2)) * 4 + t]; b1 = xf_u[(1 * 3 + (lane >> 2)) * 4 + t]; }
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(c, w0 & 0x000F000Fu, w1 & 0x000F000Fu, (w0 >> 8) & 0x000F000Fu, (w1 >> 8) & 0x000F000Fu, b0.x, b0.y);
        mma16816(c, (w0 >> 4) & 0x000F000Fu, (w1 >> 4) & 0x000F000Fu, (w0 >> 12) & 0x000F000Fu, (w1 >> 12) & 0x000F000Fu, b1.x, b1.y);
        // sum (q - 8) x = sum q x - 8 sum x: the -8 term only once (column 0 lives on the lanes with t == 0)
        const float d0 = h2f(*reinterpret_cast<const uint16_t*>(a0p + ul * 2)) * 16777216.0f;   // undo the 2^-24 of the subnormal trick
        const float d1 = h2f(*reinterpret_cast<const uint16_t*>(a1p + ul * 2)) * 16777216.0f;
        const float off = t == 0 ? 8.0f * xsm.x * (1.0f / 16777216.0f) : 0.0f;
        tot[0] += d0 * (c[0] - off); tot[1] += d0 * c[1];
        tot[2] += d1 * (c[2] - off); tot[3] += d1 * c[3];
    } else {  // Q4_K / Q5_K
        constexpr int BB = TYPE == kQ4_K ? 144 : 176;
        const int boff = (ul >> 2) * BB, gq = ul & 3;
        const uint2 W0 = *reinterpret_cast<const uint2*>(r0p + boff + 16 + gq * 32 + 8 * t);
        const uint2 W1 = *reinterpret_cast<const uint2*>(r1p + boff + 16 + gq * 32 + 8 * t);
        uint32_t h0a = 0, h0b = 0, h1a = 0, h1b = 0;   // 5th bits (Q5_K) of the two words of each row, shifted so bit 0 = low-nibble weight
        if (TYPE == kQ5_K) {
            const uint2 H0 = *reinterpret_cast<const uint2*>(r0p + boff + 144 + 8 * t);
            const uint2 H1 = *reinterpret_cast<const uint2*>(r1p + boff + 144 + 8 * t);
            h0a = H0.x >> (2 * gq); h0b = H0.y >> (2 * gq); h1a = H1.x >> (2 * gq); h1b = H1.y >> (2 * gq);
        }
        // (byte0, byte2) and (byte1, byte3) nibble pairs of word w, low (s = 0) or high (s = 4) nibbles, plus the Q5_K bit
        auto pr0 = [&](uint32_t w, uint32_t hb, int s) -> uint32_t {
            uint32_t v = (w >> s) & 0x000F000Fu;
            if (TYPE == kQ5_K) v |= ((hb >> (s ? 1 : 0)) & 0x00010001u) << 4;
            return v;
        };
        auto pr1 = [&](uint32_t w, uint32_t hb, int s) -> uint32_t {
            uint32_t v = (w >> (8 + s)) & 0x000F000Fu;
            if (TYPE == kQ5_K) v |= ((hb >> (8 + (s ? 1 : 0))) & 0x00010001u) << 4;
            return v;
        };
        uint2 bx[4];
#pragma unroll
        for (int m = 0; m < 4; m++) bx[m] = lane < 12 ? xf_u[(m * 3 + (lane >> 2)) * 4 + t] : make_uint2(0u, 0u);
        float cA[4] = {0.f, 0.f, 0.f, 0.f}, cB[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(cA, pr0(W0.x, h0a, 0), pr0(W1.x, h1a, 0), pr1(W0.x, h0a, 0), pr1(W1.x, h1a, 0), bx[0].x, bx[0].y);
        mma16816(cA, pr0(W0.y, h0b, 0), pr0(W1.y, h1b, 0), pr1(W0.y, h0b, 0), pr1(W1.y, h1b, 0), bx[1].x, bx[1].y);
        mma16816(cB, pr0(W0.x, h0a, 4), pr0(W1.x, h1a, 4), pr1(W0.x, h0a, 4), pr1(W1.x, h1a, 4), bx[2].x, bx[2].y);
        mma16816(cB, pr0(W0.y, h0b, 4), pr0(W1.y, h1b, 4), pr1(W0.y, h0b, 4), pr1(W1.y, h1b, 4), bx[3].x, bx[3].y);
        // block scales: the four lanes that share rows (g, g+8) split the decode -- t = 0: scales of row g, 1: mins of row g,
        // 2: scales of row g+8, 3: mins of row g+8 -- and trade the results by shuffle (gemv_q4k.cu:38-56 unpacking)
        const uint4 hdr = *reinterpret_cast<const uint4*>(((t & 2) ? r1p : r0p) + boff);
        const int sh = (gq & 1) * 16;
        const uint32_t hsel = (t & 1) ? hdr.z : hdr.y;   // mins live in bytes 4..7, scales in bytes 0..3 (low 6 bits); bytes 8..11 hold the high sub-blocks
        const uint32_t hlo = (hsel >> sh) & 0xFFFFu, hhi = (hdr.w >> sh) & 0xFFFFu;
        uint32_t v2;
        if (gq < 2) v2 = hlo & 0x3F3Fu;
        else v2 = (t & 1) ? (((hhi >> 4) & 0x0F0Fu) | ((hlo >> 2) & 0x3030u)) : ((hhi & 0x0F0Fu) | ((hlo >> 2) & 0x3030u));
        const float dd = (t & 1) ? h2f((uint16_t)(hdr.x >> 16)) : h2f((uint16_t)(hdr.x & 0xFFFFu)) * 16777216.0f;
        const float va = dd * (float)(v2 & 0xFFu), vb = dd * (float)(v2 >> 8);   // sub-block 2g, 2g+1
        const int base = lane & ~3;
        const float dsA0 = __shfl_sync(0xffffffffu, va, base), dsB0 = __shfl_sync(0xffffffffu, vb, base);
        const float dmA0 = __shfl_sync(0xffffffffu, va, base + 1), dmB0 = __shfl_sync(0xffffffffu, vb, base + 1);
        const float dsA1 = __shfl_sync(0xffffffffu, va, base + 2), dsB1 = __shfl_sync(0xffffffffu, vb, base + 2);
        const float dmA1 = __shfl_sync(0xffffffffu, va, base + 3), dmB1 = __shfl_sync(0xffffffffu, vb, base + 3);
        tot[0] += dsA0 * cA[0] + dsB0 * cB[0]; tot[1] += dsA0 * cA[1] + dsB0 * cB[1];
        tot[2] += dsA1 * cA[2] + dsB1 * cB[2]; tot[3] += dsA1 * cA[3] + dsB1 * cB[3];
        if (t == 0) {  // - dmin*m*sum(x): once per row (f32 sums of the unscaled x)
            tot[0] -= dmA0 * xsm.x + dmB0 * xsm.y;
            tot[2] -= dmA1 * xsm.x + dmB1 * xsm.y;
        }
    }
}

// ---- the kernel --------------------------------------------------------------
template <int TYPE, int R>
__global__ void __launch_bounds__(kSThreads, kCtasPerSm) gemv_stream_kernel(StreamW w, const SGeom g, const Prologue p, float* __restrict__ y,
                                                                            const Indirect ind) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float red[32];
    constexpr int UW = unit_w(TYPE);
    float* xs = reinterpret_cast<float*>(smem);
    float4* xsum = reinterpret_cast<float4*>(smem + g.xsum_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem + g.ring_off + (size_t)warp * g.stages * g.stage_bytes;
    const uint32_t bar0 = smem_u32(smem + g.bar_off) + warp * kMaxStages * 8;

    const int gw = blockIdx.x * kSWarps + warp, total_w = gridDim.x * kSWarps;
    const int my_tiles = gw < g.n_tiles ? (g.n_tiles - gw + total_w - 1) / total_w : 0;
    const int nq = my_tiles * g.n_slabs;
    const int groups = 32 / g.lpr;  // row groups per warp pass

    if (lane == 0) {
        for (int s = 0; s < g.stages; s++) mbar_init(bar0 + s * 8, 1);
        fence_barrier_init();
    }
    __syncwarp();

    // Issue the copies of queue entry (ti, s) into stage st.  All lanes participate.
    auto issue = [&](int ti, int s, int st) {
        const int r0 = (gw + ti * total_w) * g.rows_pass;
        const int nrows = min(g.rows_pass, w.M - r0);
        const uint32_t bar = bar0 + st * 8;
        uint8_t* sm = ring + (size_t)st * g.stage_bytes;
        if (g.contig) {  // tiles start 16-B aligned in both arrays (make_geom guarantees rows_pass*row_aux % 16 == 0)
            if (lane == 0) {
                uint32_t mb = (uint32_t)nrows * g.row_main;
                uint32_t ab = (uint32_t)((nrows * g.row_aux + 15) & ~15);
                mbar_expect_tx(bar, mb + ab);
                bulk_g2s(smem_u32(sm), w.main + (size_t)r0 * g.row_main, mb, bar);
                if (ab) bulk_g2s(smem_u32(sm + g.stage_main), w.aux + (size_t)r0 * g.row_aux, ab, bar);
            }
        } else {
            const int nun = min(g.slab_units, g.U - s * g.slab_units);
            const uint32_t mb = (uint32_t)unit_main_bytes(w.type, nun);
            const int sa = unit_aux_bytes(w.type, nun);
            const uint32_t ab = sa ? (uint32_t)((sa + 16 + 15) & ~15) : 0u;  // aligned over-fetch: row-slab scale runs start 2-B aligned
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)nrows * (mb + ab));
            __syncwarp();
            for (int rl = lane; rl < nrows; rl += 32) {
                const size_t row = (size_t)(r0 + rl);
                bulk_g2s(smem_u32(sm + rl * g.slab_main_cap), w.main + row * g.row_main + (size_t)s * g.slab_main, mb, bar);
                if (ab) {
                    uintptr_t src = reinterpret_cast<uintptr_t>(w.aux + row * g.row_aux + unit_aux_bytes(w.type, s * g.slab_units));
                    bulk_g2s(smem_u32(sm + g.stage_main + rl * g.slab_aux_cap), reinterpret_cast<const void*>(src & ~(uintptr_t)15), ab, bar);
                }
            }
        }
    };

    int ic_ti = 0, ic_s = 0, ic_st = 0, issued = 0;  // issue cursor
    auto issue_next = [&]() {
        issue(ic_ti, ic_s, ic_st);
        issued++;
        if (++ic_st == g.stages) ic_st = 0;
        if (++ic_s == g.n_slabs) { ic_s = 0; ic_ti++; }
    };

    const bool late = ind.sel != nullptr;  // expert weights are chosen by the previous kernel
    if (!late) {
        const int pre = min(nq, g.stages);
        for (int i = 0; i < pre; i++) issue_next();
    }
    pdl_launch_dependents();
    pdl_wait();
    const float* a = p.a;
    if (late) {
        const int e = ind.sel[blockIdx.y];
        y += (size_t)blockIdx.y * ind.y_stride;
        if (e < 0) {  // expert owned by another tensor-parallel rank: this slot contributes zeros to the all-reduce
            const int nout = ind.swiglu_pairs ? w.M / 2 : w.M;
            for (int i = blockIdx.x * kSThreads + threadIdx.x; i < nout; i += gridDim.x * kSThreads) y[i] = 0.0f;
            return;
        }
        w.main += (size_t)e * ind.main_stride;
        if (w.aux) w.aux += (size_t)e * ind.aux_stride;
        a += (size_t)blockIdx.y * ind.a_stride;
        const int pre = min(nq, g.stages);
        for (int i = 0; i < pre; i++) issue_next();
    }
    const unsigned int ll_epoch = ind.n_wait > 0 ? (unsigned int)(ind.wait_epoch_base[0] * ind.wait_sites_per_step + ind.wait_site + 1) : 0u;
    build_x<TYPE, kSThreads>(p, a, w.K, xs, xsum, red, blockIdx.x == 0 && blockIdx.y == 0, ll_epoch);
    constexpr bool kMma = R == 16;
    uint2* xf = reinterpret_cast<uint2*>(smem + g.xfrag_off);
    if (kMma) {
        if (TYPE == kQ4_0 || TYPE == kQ4_K || TYPE == kQ5_K) build_xfrag<TYPE>(xs, xf, w.K, 1.0f);
        __syncthreads();
    }

    if (kMma) {
        // ---- tensor-core path: the whole warp works on one unit of a 16-row tile at a time
        if (TYPE == kQ4_0 || TYPE == kQ4_K || TYPE == kQ5_K) {
            constexpr int NM = unit_w(TYPE) / 16;
            const int gq = lane >> 2, t4 = lane & 3;
            const int rstride = g.contig ? g.row_main : g.slab_main_cap, astride = g.contig ? g.row_aux : g.slab_aux_cap;
            float tot[4] = {0.f, 0.f, 0.f, 0.f};
            int ti = 0, s = 0, st = 0;
            uint32_t parity = 0;
            for (int q = 0; q < nq; q++) {
                const int r0 = (gw + ti * total_w) * 16;
                if (s == 0) tot[0] = tot[1] = tot[2] = tot[3] = 0.0f;
                const uint8_t* sm = ring + (size_t)st * g.stage_bytes;
                const uint8_t* r0p = sm + gq * rstride;
                const uint8_t* r1p = sm + (gq + 8) * rstride;
                const uint8_t* a0p = sm + g.stage_main + gq * astride;
                const uint8_t* a1p = sm + g.stage_main + (gq + 8) * astride;
                if (!g.contig && g.row_aux) {
                    a0p += reinterpret_cast<uintptr_t>(w.aux + (size_t)(r0 + gq) * g.row_aux + unit_aux_bytes(w.type, s * g.slab_units)) & 15;
                    a1p += reinterpret_cast<uintptr_t>(w.aux + (size_t)(r0 + gq + 8) * g.row_aux + unit_aux_bytes(w.type, s * g.slab_units)) & 15;
                }
                const int nun = min(g.slab_units, g.U - s * g.slab_units);
                mbar_wait(bar0 + st * 8, parity);
                for (int ul = 0; ul < nun; ul++) {
                    const int u = s * g.slab_units + ul;
                    float4 xsm = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (TYPE == kQ4_K || TYPE == kQ5_K) xsm = xsum[u];
                    if (TYPE == kQ4_0) {  // sum of the unit's 32 x (the -8 offset term)
                        const float4* xp = reinterpret_cast<const float4*>(xs + u * 36);
                        float sx = 0.0f;
#pragma unroll
                        for (int j = 0; j < 8; j++) { float4 v4 = xp[j]; sx += (v4.x + v4.y) + (v4.z + v4.w); }
                        xsm.x = sx;
                    }
                    unit_mma<TYPE>(r0p, r1p, a0p, a1p, ul, xf + (size_t)u * NM * 12, xsm, tot, lane);
                }
                __syncwarp();
                if (issued < nq) {
                    fence_proxy_async();
                    issue_next();
                }
                if (s == g.n_slabs - 1) {
                    // columns 0..2 hold the three fp16 terms of x: add them (they sit on the lanes t = 0, 1 of each row group)
                    float vlo = tot[0] + tot[1], vhi = tot[2] + tot[3];
                    vlo += __shfl_xor_sync(0xffffffffu, vlo, 1); vhi += __shfl_xor_sync(0xffffffffu, vhi, 1);
                    vlo += __shfl_xor_sync(0xffffffffu, vlo, 2); vhi += __shfl_xor_sync(0xffffffffu, vhi, 2);
                    const int rlo = r0 + gq, rhi = r0 + gq + 8;
                    if (ind.n_peers > 0) {
                        if (t4 == 0) {
                            const unsigned int epoch = (unsigned int)(ind.epoch_base[0] * ind.sites_per_step + ind.site + 1);
                            for (int pp = 0; pp < ind.n_peers; pp++) {
                                if (rlo < w.M)
                                    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2*>(ind.peer_out[pp]) + rlo),
                                                 "r"(__float_as_uint(vlo)), "r"(epoch) : "memory");
                                if (rhi < w.M)
                                    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2*>(ind.peer_out[pp]) + rhi),
                                                 "r"(__float_as_uint(vhi)), "r"(epoch) : "memory");
                            }
                        }
                    } else if (!ind.swiglu_pairs) {
                        if (t4 == 0) {
                            if (rlo < w.M) y[rlo] = vlo;
                            if (rhi < w.M) y[rhi] = vhi;
                        }
                    } else {  // (gate, up) = rows (2i, 2i+1): the up row lives 4 lanes further
                        const float ulo = __shfl_down_sync(0xffffffffu, vlo, 4), uhi = __shfl_down_sync(0xffffffffu, vhi, 4);
                        if (t4 == 0 && !(gq & 1)) {
                            if (rlo + 1 < w.M) { double gv = (double)vlo; y[rlo >> 1] = (float)(gv * (1.0 / (1.0 + exp(-gv)))) * ulo; }
                            if (rhi + 1 < w.M) { double gv = (double)vhi; y[rhi >> 1] = (float)(gv * (1.0 / (1.0 + exp(-gv)))) * uhi; }
                        }
                    }
                }
                if (++s == g.n_slabs) { s = 0; ti++; }
                if (++st == g.stages) { st = 0; parity ^= 1u; }
            }
        }
        return;
    }
    const int sr = lane / g.lpr, lr = lane % g.lpr;
    // per-lane row offsets inside a stage do not change from stage to stage
    int offm[R], offa[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int rl = j * groups + sr;
        offm[j] = rl * (g.contig ? g.row_main : g.slab_main_cap);
        offa[j] = g.stage_main + rl * (g.contig ? g.row_aux : g.slab_aux_cap);
    }
    float acc[R];
    int ti = 0, s = 0, st = 0;
    uint32_t parity = 0;
    for (int q = 0; q < nq; q++) {
        const int r0 = (gw + ti * total_w) * g.rows_pass;
        if (s == 0) {
#pragma unroll
            for (int j = 0; j < R; j++) acc[j] = 0.0f;
        }
        const uint8_t* sm = ring + (size_t)st * g.stage_bytes;
        const uint8_t* rowm[R];
        const uint8_t* rowa[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            rowm[j] = sm + offm[j];
            rowa[j] = sm + offa[j];
            if (!g.contig && g.row_aux) {  // slab mode: the scale run of this row-slab was fetched from its 16-B aligned floor
                const int rl = j * groups + sr;
                uintptr_t src = reinterpret_cast<uintptr_t>(w.aux + (size_t)(r0 + rl) * g.row_aux + unit_aux_bytes(w.type, s * g.slab_units));
                rowa[j] += (src & 15);
            }
        }
        const int nun = min(g.slab_units, g.U - s * g.slab_units);
        mbar_wait(bar0 + st * 8, parity);
        for (int i = 0; i < g.upl; i++) {
            const int ul = lr + g.lpr * i;
            if (ul >= nun) break;
            const int u = s * g.slab_units + ul;
            float4 xsm = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (TYPE == kQ4_K || TYPE == kQ5_K) xsm = xsum[u];
            unit_dot<TYPE, R>(rowm, rowa, ul, xs + u * (UW + 4), 0, xsm, acc, lane);
        }
        __syncwarp();
        if (issued < nq) {  // refill the stage just drained (generic-proxy reads ordered before the async-proxy write)
            fence_proxy_async();
            issue_next();
        }
        if (s == g.n_slabs - 1) {
            float v[R];
#pragma unroll
            for (int j = 0; j < R; j++) {
                v[j] = acc[j];
                for (int o = g.lpr >> 1; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
            }
            if (ind.n_peers > 0) {  // row-parallel partial sums: push every row into this rank's slot on every peer
#pragma unroll
                for (int j = 0; j < R; j++) {
                    const int row = r0 + j * groups + sr;
                    if (lr == 0 && row < w.M) {  // 8-byte (value, epoch) stores are delivered atomically: the data is its own flag (LL protocol)
                        const unsigned int epoch = (unsigned int)(ind.epoch_base[0] * ind.sites_per_step + ind.site + 1);
                        for (int pp = 0; pp < ind.n_peers; pp++)
                            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2*>(ind.peer_out[pp]) + row),
                                         "r"(__float_as_uint(v[j])), "r"(epoch)
                                         : "memory");
                    }
                }
            } else if (!ind.swiglu_pairs) {
#pragma unroll
                for (int j = 0; j < R; j++) {
                    const int row = r0 + j * groups + sr;
                    if (lr == 0 && row < w.M) y[row] = v[j];  // rows past M (ragged last tile) were computed on stale smem and are dropped
                }
            } else if (groups == 1) {  // rows j, j+1 of the pair live in the same lanes (R is even here)
#pragma unroll
                for (int j = 0; j + 1 < R; j += 2) {
                    const int row = r0 + j;
                    if (lr == 0 && row + 1 < w.M) {
                        double gv = (double)v[j];
                        y[row >> 1] = (float)(gv * (1.0 / (1.0 + exp(-gv)))) * v[j + 1];
                    }
                }
            } else {                   // the up row sits in the next lane group
#pragma unroll
                for (int j = 0; j < R; j++) {
                    float up = __shfl_down_sync(0xffffffffu, v[j], g.lpr);
                    const int row = r0 + j * groups + sr;
                    if (lr == 0 && !(sr & 1) && row + 1 < w.M) {
                        double gv = (double)v[j];
                        y[row >> 1] = (float)(gv * (1.0 / (1.0 + exp(-gv)))) * up;
                    }
                }
            }
        }
        if (++s == g.n_slabs) { s = 0; ti++; }
        if (++st == g.stages) { st = 0; parity ^= 1u; }
    }
}

// ---- host side -----------------------------------------------------------------
int pick_lpr(int U) {
    int best = 32;
    double best_eff = 0.0;
    const int cand[4] = {32, 16, 8, 4};
    for (int i = 0; i < 4; i++) {
        int l = cand[i];
        int ct = (U + l - 1) / l;
        double eff = (double)U / ((double)ct * l);
        if (eff >= 0.9) return l;  // largest lane count that wastes < 10 %
        if (eff > best_eff) { best_eff = eff; best = l; }
    }
    return best;
}

// Returns false when the shape does not fit the requested tile mode / occupancy.
bool make_geom(int type, int M, int K, int R, bool want_contig, int ctas_per_sm, bool pairs, SGeom& g, bool mma = false) {
    const bool kq = stream_is_kquant(type);
    if (K % (kq ? 256 : 32) || M <= 0 || stream_main_per8(type) == 0) return false;
    if (kq && R > 2 && !mma) return false;  // K-quants: the unrolled low/high-nibble body of R > 2 rows overflows registers and the L0 I-cache
    const int uw = unit_w(type);
    g.U = K / uw;
    g.mma = mma ? 1 : 0;
    if (mma && !(type == kQ4_0 || type == kQ4_K || type == kQ5_K)) return false;
    g.lpr = mma ? 32 : pick_lpr(g.U);
    g.rows_pass = mma ? 16 : (32 / g.lpr) * R;
    if (pairs && ((g.rows_pass & 1) || (M & 1))) return false;  // (gate_i, up_i) row pairs must not straddle tiles
    g.row_main = unit_main_bytes(type, g.U);
    g.row_aux = unit_aux_bytes(type, g.U);
    g.ctas_per_sm = ctas_per_sm;
    const int budget = kSmemTotal / ctas_per_sm - 1024;
    const int xbytes = ((g.U * (uw + 4) * 4 + 127) & ~127);
    const int xsum_bytes = (type == kQ4_K || type == kQ5_K) ? ((g.U * 16 + 127) & ~127) : 0;
    const int xfrag_bytes = mma ? ((g.U * (uw / 16) * 96 + 127) & ~127) : 0;  // fp16 B fragments: 3 splits x 4 lanes x 8 B per MMA
    int ring_budget = budget - xbytes - xsum_bytes - xfrag_bytes - 512;
    if (ring_budget < 8 * 1024) return false;
    const int warp_budget = (ring_budget / kSWarps) & ~15;
    const int ct = mma ? g.U : (g.U + g.lpr - 1) / g.lpr;  // units per lane per row (tensor-core path: every lane walks every unit)
    const int lpr_eff = mma ? 1 : g.lpr;
    if (want_contig) {
        if (g.row_aux && (g.rows_pass * g.row_aux) % 16) return false;  // every tile must start 16-B aligned in the scale array
        int tile_main = g.rows_pass * g.row_main;
        int tile_aux = (g.rows_pass * g.row_aux + 15) & ~15;
        if (2 * (tile_main + tile_aux) > warp_budget) return false;
        g.contig = 1;
        g.upl = ct;
        g.slab_units = lpr_eff * ct;
        g.n_slabs = 1;
        g.slab_main = g.row_main;
        g.slab_main_cap = g.row_main;
        g.slab_aux_cap = 0;
        g.stage_main = tile_main;
        g.stage_bytes = tile_main + tile_aux;
    } else {
        // K-slabs: the largest whole number of units per lane such that >= 2 stages of >= 2 KB... fit
        const int unit = kq ? 4 : 1;  // slabs must hold whole super-blocks
        int upl = 0;
        for (int c = ct; c >= 1; c--) {
            int su = lpr_eff * c;
            if (su % unit) continue;
            int sm_ = unit_main_bytes(type, su);
            int sa = unit_aux_bytes(type, su);
            int sac = sa ? ((sa + 16 + 15) & ~15) : 0;
            if (3 * g.rows_pass * (sm_ + sac) <= warp_budget) { upl = c; break; }
        }
        if (!upl) return false;
        g.contig = 0;
        g.upl = upl;
        g.slab_units = lpr_eff * upl;
        g.n_slabs = (g.U + g.slab_units - 1) / g.slab_units;
        g.slab_main = unit_main_bytes(type, g.slab_units);
        g.slab_main_cap = g.slab_main;
        int sa = unit_aux_bytes(type, g.slab_units);
        g.slab_aux_cap = sa ? ((sa + 16 + 15) & ~15) : 0;
        g.stage_main = g.rows_pass * g.slab_main_cap;
        g.stage_bytes = g.stage_main + g.rows_pass * g.slab_aux_cap;
    }
    g.stages = warp_budget / g.stage_bytes;
    if (g.stages > kMaxStages) g.stages = kMaxStages;
    if (g.stages < 2) return false;
    g.n_tiles = (M + g.rows_pass - 1) / g.rows_pass;
    g.xsum_off = xbytes;
    g.xfrag_off = xbytes + xsum_bytes;
    g.ring_off = xbytes + xsum_bytes + xfrag_bytes;
    g.bar_off = (g.ring_off + kSWarps * g.stages * g.stage_bytes + 15) & ~15;
    g.smem_bytes = g.bar_off + kSWarps * kMaxStages * 8;
    return g.smem_bytes <= budget;
}

// Tile shape policy: the most co-resident CTAs per SM first (thread-level parallelism hides the dequant latency),
// then whole-row contiguous tiles with the most rows per x load, K-slabs for long rows; among the feasible shapes
// the first that still gives every warp slot of the chip a tile wins, else the one with the smallest tiles.
int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && v[0]) ? atoi(v) : dflt;
}

bool choose_geom(int type, int M, int K, bool pairs, SGeom& best, int& bestR) {
    static const struct { int R; bool contig; } order[6] = {{4, true}, {2, true}, {4, false}, {1, true}, {2, false}, {1, false}};
    static const int force_cps = env_int("ZB_GEMV_CPS", 0), force_r = env_int("ZB_GEMV_R", 0);  // tuning knobs (experiments only)
    bestR = 0;
    static const int use_mma = env_int("ZB_GEMV_MMA", 0);  // tensor-core dequant path (Q4_0, Q4_K, Q5_K)
    if (use_mma && (type == kQ4_0 || type == kQ4_K || type == kQ5_K)) {
        const int min_rows = env_int("ZB_GEMV_MMA_MIN_ROWS", 0);
        if (M >= min_rows)
            for (int cps = kCtasPerSm; cps >= 1; cps--) {
                SGeom g{};
                if (make_geom(type, M, K, 16, true, cps, pairs, g, true) || make_geom(type, M, K, 16, false, cps, pairs, g, true)) {
                    best = g;
                    bestR = 16;
                    return true;
                }
            }
    }
    if (force_cps || force_r) {
        for (int cps = force_cps ? force_cps : kCtasPerSm; cps >= 1 && !bestR; cps--) {
            SGeom g{};
            for (int i = 0; i < 6 && !bestR; i++) {
                if (force_r && order[i].R != force_r) continue;
                if (make_geom(type, M, K, order[i].R, order[i].contig, cps, pairs, g)) { best = g; bestR = order[i].R; }
            }
            if (force_cps) break;
        }
        if (bestR) return true;
    }
    // Streaming-size matrices (>= 24 MB) run 2 CTAs per SM: the larger rings buy rows-per-x-load (measured: lm_head 163 -> 112 us);
    // everything smaller is latency-bound and wants the 24 warps per SM of 3 CTAs.
    const long long bytes = (long long)M * unit_main_bytes(type, K / unit_w(type));
    const int cps_order[3] = {bytes >= (24ll << 20) ? 2 : 3, bytes >= (24ll << 20) ? 3 : 2, 1};
    for (int ci = 0; ci < 3 && !bestR; ci++) {
        const int cps = cps_order[ci];
        const int slots = ZB_SMS * kSWarps * cps;
        SGeom g{};
        for (int i = 0; i < 6; i++) {
            if (!make_geom(type, M, K, order[i].R, order[i].contig, cps, pairs, g)) continue;
            if (!bestR || g.rows_pass < best.rows_pass) { best = g; bestR = order[i].R; }
            if (g.n_tiles >= slots) { best = g; bestR = order[i].R; break; }
        }
    }
    return bestR != 0;
}

template <int TYPE, int R>
cudaError_t launch_t(const StreamW& w, const SGeom& g, const Prologue& p, float* y, const Indirect& ind, int nsel, bool pdl, cudaStream_t stream) {
    static DeviceOnce once;
    if (cudaError_t e = once.ensure(1, [] { return cudaFuncSetAttribute(gemv_stream_kernel<TYPE, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal - 2048); }))
        return e;
    int ctas = (g.n_tiles + kSWarps - 1) / kSWarps;
    // All CTA slots are used: leaving one slot per SM to the next kernel of the PDL chain (so that it prefetches while this
    // one computes) was measured slower (C2 538 vs 578 tok/s) -- the lost warps cost more than the hidden prologue.
    static const int grid_cps = env_int("ZB_GEMV_GRID_CPS", 0);
    int use = grid_cps > 0 ? grid_cps : g.ctas_per_sm;
    if (use > g.ctas_per_sm) use = g.ctas_per_sm;
    int cap = ZB_SMS * use;
    if (nsel > 1) cap = (cap + nsel - 1) / nsel;
    if (ctas > cap) ctas = cap;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas, nsel > 0 ? nsel : 1, 1);
    cfg.blockDim = dim3(kSThreads, 1, 1);
    cfg.dynamicSmemBytes = g.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, gemv_stream_kernel<TYPE, R>, w, g, p, y, ind);
}

template <int TYPE>
cudaError_t launch_r(const StreamW& w, const Prologue& p, float* y, const Indirect& ind, int nsel, bool pdl, cudaStream_t stream) {
    SGeom best{};
    int bestR = 0;
    if (!choose_geom(TYPE, w.M, w.K, ind.swiglu_pairs != 0, best, bestR)) return cudaErrorInvalidConfiguration;
    switch (bestR) {
        case 16: return launch_t<TYPE, 16>(w, best, p, y, ind, nsel, pdl, stream);
        case 4: return launch_t<TYPE, 4>(w, best, p, y, ind, nsel, pdl, stream);
        case 2: return launch_t<TYPE, 2>(w, best, p, y, ind, nsel, pdl, stream);
        default: return launch_t<TYPE, 1>(w, best, p, y, ind, nsel, pdl, stream);
    }
}

cudaError_t launch_any(const StreamW& w, const Prologue& p, float* y, const Indirect& ind, int nsel, bool pdl, cudaStream_t stream) {
    switch (w.type) {
        case kQ4_0: return launch_r<kQ4_0>(w, p, y, ind, nsel, pdl, stream);
        case kQ8_0: return launch_r<kQ8_0>(w, p, y, ind, nsel, pdl, stream);
        case kQ4_K: return launch_r<kQ4_K>(w, p, y, ind, nsel, pdl, stream);
        case kQ5_K: return launch_r<kQ5_K>(w, p, y, ind, nsel, pdl, stream);
        case kQ6_K: return launch_r<kQ6_K>(w, p, y, ind, nsel, pdl, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

// ===========================================================================
// C ABI (include/zb200.h)
// ===========================================================================
ZB_API int zb_stream_layout(int qtype, int rows, int cols, int64_t* main_bytes, int64_t* aux_bytes) {
    const bool kq = zb::stream_is_kquant(qtype);
    if (zb::stream_main_per8(qtype) == 0 || cols % (kq ? 256 : 32) || rows < 0) return cudaErrorInvalidValue;
    if (main_bytes) *main_bytes = (int64_t)rows * zb::stream_main_bytes(qtype, cols / 32);
    if (aux_bytes) {
        int64_t a = (int64_t)rows * zb::stream_aux_bytes(qtype, cols / 32);
        *aux_bytes = a ? ((a + 15) & ~(int64_t)15) + 32 : 0;  // padded: aux copies are 16-B aligned over-fetches
    }
    return 0;
}

// Host-side repack of raw GGUF blocks into the stream layout (pure byte moves).
ZB_API int zb_stream_repack_host(int qtype, const void* raw, int rows, int cols, void* main_out, void* aux_out) {
    const uint8_t* src = static_cast<const uint8_t*>(raw);
    uint8_t* m = static_cast<uint8_t*>(main_out);
    uint8_t* a = static_cast<uint8_t*>(aux_out);
    const int64_t n32 = (int64_t)rows * (cols / 32), n256 = (int64_t)rows * (cols / 256);
    constexpr int64_t kChunk = 1 << 14;   // blocks per work item of the host thread pool
    auto chunks = [&](int64_t n, auto fn) {
        zb::zb_parallel_for((n + kChunk - 1) / kChunk, 4, [&](int64_t c) {
            const int64_t lo = c * kChunk, hi = lo + kChunk < n ? lo + kChunk : n;
            fn(lo, hi);
        });
    };
    switch (qtype) {
        case zb::kQ4_0:
            chunks(n32, [&](int64_t lo, int64_t hi) { for (int64_t b = lo; b < hi; b++) { memcpy(a + b * 2, src + b * 18, 2); memcpy(m + b * 16, src + b * 18 + 2, 16); } });
            return 0;
        case zb::kQ8_0:
            chunks(n32, [&](int64_t lo, int64_t hi) { for (int64_t b = lo; b < hi; b++) { memcpy(a + b * 2, src + b * 34, 2); memcpy(m + b * 32, src + b * 34 + 2, 32); } });
            return 0;
        case zb::kQ4_K: chunks(n256, [&](int64_t lo, int64_t hi) { memcpy(m + lo * 144, src + lo * 144, (size_t)(hi - lo) * 144); }); return 0;
        case zb::kQ5_K: chunks(n256, [&](int64_t lo, int64_t hi) { memcpy(m + lo * 176, src + lo * 176, (size_t)(hi - lo) * 176); }); return 0;
        case zb::kQ6_K:
            chunks(n256, [&](int64_t lo, int64_t hi) { for (int64_t b = lo; b < hi; b++) { memcpy(m + b * 208, src + b * 210, 208); memcpy(a + b * 2, src + b * 210 + 208, 2); } });
            return 0;
    }
    return cudaErrorInvalidValue;
}

// 0 when a [rows, cols] matrix of this type can be run by the streamed kernel.
ZB_API int zb_stream_check(int qtype, int rows, int cols) {
    if (zb::stream_main_per8(qtype) == 0) return cudaErrorInvalidValue;
    SGeom g{};
    int R = 0;
    return choose_geom(qtype, rows, cols, false, g, R) ? 0 : (int)cudaErrorInvalidConfiguration;
}

ZB_API int zb_gemv_stream_f32(const zb_stream_weight* w, const zb_prologue* p, float* y, int flags, zb_stream_t stream) {
    if (!w || !p || !y) return cudaErrorInvalidValue;
    zb::StreamW sw{static_cast<const uint8_t*>(w->main), static_cast<const uint8_t*>(w->aux), w->qtype, w->rows, w->cols};
    zb::Prologue pr{p->a, p->r, p->w1, p->w2, p->sum_out, p->mix_w, p->mix_n, p->mix_stride, p->eps, p->swiglu};
    Indirect ind{};
    int nsel = 0;
    if (w->expert_sel) {
        ind.sel = w->expert_sel;
        ind.main_stride = w->expert_main_stride;
        ind.aux_stride = w->expert_aux_stride;
        ind.a_stride = p->a_slot_stride;
        ind.y_stride = w->y_slot_stride;
        nsel = w->n_sel;
        if (nsel <= 0) return cudaErrorInvalidValue;
    }
    ind.swiglu_pairs = w->epilogue == 1;
    if (w->n_peers > 0) {
        if (w->n_peers > 8 || !w->epoch_base || w->expert_sel) return cudaErrorInvalidValue;
        ind.n_peers = w->n_peers; ind.site = w->site; ind.sites_per_step = w->sites_per_step;
        for (int i = 0; i < w->n_peers; i++) { ind.peer_out[i] = w->peer_out[i]; ind.peer_flag[i] = w->peer_flag[i]; }
        ind.epoch_base = w->epoch_base;
        ind.ticket = w->ticket;
    }
    if (p->n_wait > 0) {
        if (p->n_wait > 8 || !p->wait_epoch_base) return cudaErrorInvalidValue;
        ind.wait_flags = p->wait_flags; ind.wait_epoch_base = p->wait_epoch_base;
        ind.n_wait = p->n_wait; ind.wait_site = p->wait_site; ind.wait_sites_per_step = p->wait_sites_per_step;
    }
    return launch_any(sw, pr, y, ind, nsel, (flags & 1) != 0, (cudaStream_t)stream);
}
