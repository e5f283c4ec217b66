// Tensor-core fused dequant-GEMV for batch-1 decode (sm_100a): y[M] = deq(W[M,K]) . x[K], x = prologue(...).
//
// Why tensor cores at batch 1.  At B200 rates (6.5 TB/s over 148 SMs) a 4-bit GEMV has an issue budget of ~3.2 lane
// instructions per weight; the CUDA-core dequant of gemv_stream.cu needs ~3.8 (PRMT + FADD2 + FFMA2 per weight) and tops
// out near half the HBM roofline.  Here the int->float conversion AND the multiply-accumulate run on the tensor pipe:
//   * a nibble pair masked out of a quant word IS an fp16x2 operand -- the subnormals n * 2^-24 (low nibbles) and
//     n * 2^-20 (high nibbles), exact -- so a weight costs ~0.6 ALU instructions (one LOP3 per two weights);
//   * the activation vector enters as three fp16 terms x*s = h1 + h2 + h3 (s a power of two chosen from max|x|, ~33
//     significant bits) in three of the eight B columns of mma.sync.m16n8k16; products are exact, accumulation is f32;
//   * block scales are applied to the f32 accumulators per 32-weight sub-block (two MMAs), the Q4_K min term
//     sum_s m_s * sum(x_s) is itself one more MMA per super-block (6-bit mins as subnormals x per-sub-block sums of x).
// Work unit: a "block-tile" = 16 rows x one 256-weight super-block (2304 B for Q4_K), laid out at upload time so that
// every lane's operand bytes are one conflict-free LDS.128 (pure byte permutation of the GGUF blocks: dequantised values
// stay bit-exact).  Block-tiles are numbered (row_tile * K/256 + super_block) = their order in memory; CTA c owns the
// contiguous range [c*q, (c+1)*q), warp w of the CTA takes every 16th of them through a private TMA ring
// (cp.async.bulk + mbarrier, first fill issued before griddepcontrol.wait so it overlaps the previous kernel).  Row tiles
// that straddle CTAs are finished by the last CTA to arrive (atomic ticket, fixed summation order: deterministic).
// One CTA of 16 warps per SM: the fused prologue (build_x, zb_prologue.cuh) runs once per SM instead of once per 8 warps.
//
// Reference semantics replaced: Engine.MatMul on Q4_K storage (gemv_q4k.cu:68-160, dequant spec :14-21,38-56) plus the
// fused providers around it (fused_add_rmsnorm.cu:17-84, fused_norm_add.cu:11-81, fused_swiglu.cu:11-34).
#include <stdlib.h>
#include <string.h>

#include "zb_prologue.cuh"
#include "zb200.h"

namespace {

using namespace zb;

constexpr int kMW = 16;                  // warps per CTA
constexpr int kMT = kMW * 32;
constexpr int kMStagesMax = 4;
constexpr int kMSmem = 225 * 1024;
constexpr int kMaxParts = 8;             // CTAs that may share one row tile

__host__ __device__ constexpr int bt_bytes(int type) { return type == kQ4_K ? 2304 : (type == kQ6_K ? 3360 : 0); }
constexpr int kXmWords = 16;             // per super-block: Q4_K 12 half2 min-term fragments, Q6_K 16 f32 offset terms

struct MGeom {
    int nb;            // super-blocks per row
    int n_tiles;       // 16-row tiles
    int total;         // block-tiles
    int per_cta, ctas;
    int stages, slots, max_local;
    int xsum_off, xf_off, xm_off, part_off, ring_off, bar_off, smem_bytes;
};

__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : 0.0f;
    return warp_max(t);
}

// D = A(16x16, row) * B(16x8, col) + C, f16 operands, f32 accumulate
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1,
                                        const float (&c)[4]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

__device__ __forceinline__ float2 h2x2_to_f2(uint32_t v) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    return __half22float2(h);
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// x (unit-major f32, 64-weight units padded to 68) -> fp16 B fragments in the k order the nibble pairs come out in.
// One thread per (super-block b, group G, nibble plane, t); a warp covers one super-block.
//   xf[((b*8 + G*2 + nib)*12 + n*4 + t)] (uint4) = {b0,b1 of MMA j=0, b0,b1 of MMA j=1} of split term n
//   xm[b*16 + n*4 + G] (uint32)                   = half2(term_n(sum x of sub-block 2G), term_n(sum x of sub-block 2G+1))
__device__ void build_frags_q4k(const float* xs, const float4* xsum, uint4* xf, uint32_t* xm, int nb, float s) {
    const int lane = threadIdx.x & 31;
    for (int b = threadIdx.x >> 5; b < nb; b += kMW) {
        const int t = lane & 3, nib = (lane >> 2) & 1, G = lane >> 3;
        const int u = b * 4 + G;
        const float sc = nib ? s * 0.0625f : s;   // high nibbles enter the MMA as n * 2^-20: their x carries the 2^-4
        const float4* xp = reinterpret_cast<const float4*>(xs + u * 68 + nib * 32 + 8 * t);
        const float4 v0 = xp[0], v1 = xp[1];
        float v[8] = {v0.x * sc, v0.y * sc, v0.z * sc, v0.w * sc, v1.x * sc, v1.y * sc, v1.z * sc, v1.w * sc};
        const float4 su = xsum[u];
        float sv = (nib ? su.y : su.x) * s * 0.015625f;   // sum of 32 x: 2^-6 keeps it inside fp16 range
#pragma unroll
        for (int n = 0; n < 3; n++) {
            __half h[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                h[e] = __float2half_rn(v[e]);
                v[e] -= __half2float(h[e]);
            }
            uint4 o;
            o.x = pack_h2(h[0], h[2]);
            o.y = pack_h2(h[1], h[3]);
            o.z = pack_h2(h[4], h[6]);
            o.w = pack_h2(h[5], h[7]);
            xf[(size_t)((b * 8 + G * 2 + nib) * 12 + n * 4 + t)] = o;
            const __half hs = __float2half_rn(sv);
            sv -= __half2float(hs);
            const uint32_t mine = (uint32_t)__half_as_ushort(hs);
            const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 4);
            if (nib == 0 && t == 0) xm[b * kXmWords + n * 4 + G] = mine | (other << 16);
        }
    }
}

// One Q4_K block-tile (16 rows x 256 weights) on the tensor cores.  tot[0..1] += row g, columns (2t, 2t+1);
// tot[2..3] += row g+8.  Columns 0..2 are the three fp16 terms of x (the other columns repeat them and are ignored).
__device__ __forceinline__ void block_tile_q4k(const uint8_t* bt, const uint4* xfb, const uint32_t* xmb, float (&tot)[4], int lane,
                                               int bsel, uint32_t msel) {
    const int g = lane >> 2;
    const uint4* q = reinterpret_cast<const uint4*>(bt);
    uint4 qa[2], qb[2];
    qa[0] = q[lane]; qa[1] = q[32 + lane];          // row g:   groups (0,1), (2,3)
    qb[0] = q[64 + lane]; qb[1] = q[96 + lane];     // row g+8
    const uint4 ha = q[128 + g], hb = q[136 + g];   // d | dmin | 12 packed scale bytes (gemv_q4k.cu:38-56)
    // 6-bit scales / mins of the 8 sub-blocks, four to a word
    uint32_t sca[2], scb[2], mna[2], mnb[2];
    sca[0] = ha.y & 0x3F3F3F3Fu; sca[1] = (ha.w & 0x0F0F0F0Fu) | ((ha.y >> 2) & 0x30303030u);
    scb[0] = hb.y & 0x3F3F3F3Fu; scb[1] = (hb.w & 0x0F0F0F0Fu) | ((hb.y >> 2) & 0x30303030u);
    mna[0] = ha.z & 0x3F3F3F3Fu; mna[1] = ((ha.w >> 4) & 0x0F0F0F0Fu) | ((ha.z >> 2) & 0x30303030u);
    mnb[0] = hb.z & 0x3F3F3F3Fu; mnb[1] = ((hb.w >> 4) & 0x0F0F0F0Fu) | ((hb.z >> 2) & 0x30303030u);
    const float zero[4] = {0.f, 0.f, 0.f, 0.f};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int G = 0; G < 4; G++) {
        const uint32_t wa0 = (G & 1) ? qa[G >> 1].z : qa[G >> 1].x, wa1 = (G & 1) ? qa[G >> 1].w : qa[G >> 1].y;
        const uint32_t wb0 = (G & 1) ? qb[G >> 1].z : qb[G >> 1].x, wb1 = (G & 1) ? qb[G >> 1].w : qb[G >> 1].y;
        const uint4 bl = xfb[(G * 2) * 12 + bsel], bh = xfb[(G * 2 + 1) * 12 + bsel];
        const uint32_t sa0 = wa0 >> 8, sa1 = wa1 >> 8, sb0 = wb0 >> 8, sb1 = wb1 >> 8;
        float cl[4], ch[4];
        mma_f16(cl, wa0 & 0x000F000Fu, wb0 & 0x000F000Fu, sa0 & 0x000F000Fu, sb0 & 0x000F000Fu, bl.x, bl.y, zero);
        mma_f16(cl, wa1 & 0x000F000Fu, wb1 & 0x000F000Fu, sa1 & 0x000F000Fu, sb1 & 0x000F000Fu, bl.z, bl.w, cl);
        mma_f16(ch, wa0 & 0x00F000F0u, wb0 & 0x00F000F0u, sa0 & 0x00F000F0u, sb0 & 0x00F000F0u, bh.x, bh.y, zero);
        mma_f16(ch, wa1 & 0x00F000F0u, wb1 & 0x00F000F0u, sa1 & 0x00F000F0u, sb1 & 0x00F000F0u, bh.z, bh.w, ch);
        // scales of sub-blocks (2G, 2G+1) as fp16 subnormals sc * 2^-24 -> f32
        const uint32_t sel = (G & 1) ? 0x4342u : 0x4140u;
        const float2 fa = h2x2_to_f2(__byte_perm(sca[G >> 1], 0u, sel));
        const float2 fb = h2x2_to_f2(__byte_perm(scb[G >> 1], 0u, sel));
        acc[0] = fmaf(fa.x, cl[0], acc[0]); acc[1] = fmaf(fa.x, cl[1], acc[1]);
        acc[2] = fmaf(fb.x, cl[2], acc[2]); acc[3] = fmaf(fb.x, cl[3], acc[3]);
        acc[0] = fmaf(fa.y, ch[0], acc[0]); acc[1] = fmaf(fa.y, ch[1], acc[1]);
        acc[2] = fmaf(fb.y, ch[2], acc[2]); acc[3] = fmaf(fb.y, ch[3], acc[3]);
    }
    // min term: A[row][k = sub-block] = m * 2^-24 (k 8..15 zero), B[k][n] = term_n(sum x of sub-block k) * 2^-6
    const int t = lane & 3;
    const uint32_t ma = __byte_perm(t < 2 ? mna[0] : mna[1], 0u, msel);
    const uint32_t mb = __byte_perm(t < 2 ? mnb[0] : mnb[1], 0u, msel);
    float cm[4];
    mma_f16(cm, ma, mb, 0u, 0u, xmb[bsel], 0u, zero);
    const float da = h2f((uint16_t)(ha.x & 0xFFFFu)), dma = h2f((uint16_t)(ha.x >> 16)) * 3.814697265625e-06f;  // 2^-18
    const float db = h2f((uint16_t)(hb.x & 0xFFFFu)), dmb = h2f((uint16_t)(hb.x >> 16)) * 3.814697265625e-06f;
    tot[0] = fmaf(-dma, cm[0], fmaf(da, acc[0], tot[0])); tot[1] = fmaf(-dma, cm[1], fmaf(da, acc[1], tot[1]));
    tot[2] = fmaf(-dmb, cm[2], fmaf(db, acc[2], tot[2])); tot[3] = fmaf(-dmb, cm[3], fmaf(db, acc[3], tot[3]));
}


// ---- Q6_K (gemv_q6k.cu:11-25): 16 scale groups of 16 consecutive weights per super-block, one MMA each ----------------
// Block-tile (3360 B = 16 x 210): [row half h][ql run A | ql run B | qh][lane][16 B], then int8 scales [h][g][16], then fp16 d.
// Lane (g, t) owns, per (half hf, is): the ql words A = ql[64hf + 16is + 4t ..], B = ql[64hf + 32 + 16is + 4t ..] and the qh
// word qh[32hf + 16is + 4t ..]: low nibbles + qh bits (0-1 | 2-3) are q1 | q2, high nibbles + bits (4-5 | 6-7) are q3 | q4.
//   xf as uint2[((b*8 + sg/2)*12 + n*4 + t)*2 + (sg&1)] = (b0, b1) of scale group sg, split term n
//   xm as float[b*16 + sg] = -32 * 2^-24 * s * sum(x of group sg): the "- 32" of every weight, fed in as the MMA's C operand
__device__ void build_frags_q6k(const float* xs, uint2* xf2, float* off, int nb, float s) {
    for (int idx = threadIdx.x; idx < nb * 64; idx += kMT) {
        const int t = idx & 3, sg = (idx >> 2) & 15, b = idx >> 6;
        const int e = b * 256 + sg * 16 + 4 * t;
        const float4 v4 = *reinterpret_cast<const float4*>(xs + (e >> 6) * 68 + (e & 63));
        const bool hi = ((sg >> 1) & 3) >= 2;            // q3 | q4 enter as n * 2^-20
        const float sc = hi ? s * 0.0625f : s;
        float v[4] = {v4.x * sc, v4.y * sc, v4.z * sc, v4.w * sc};
        float sum = (v4.x + v4.y) + (v4.z + v4.w);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (t == 0) off[b * 16 + sg] = -1.9073486328125e-06f * s * sum;   // 32 * 2^-24 = 2^-19
#pragma unroll
        for (int n = 0; n < 3; n++) {
            __half h[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                h[k] = __float2half_rn(v[k]);
                v[k] -= __half2float(h[k]);
            }
            xf2[(size_t)(((b * 8 + (sg >> 1)) * 12 + n * 4 + t) * 2 + (sg & 1))] = make_uint2(pack_h2(h[0], h[2]), pack_h2(h[1], h[3]));
        }
    }
}

__device__ __forceinline__ uint32_t word_of(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
__device__ __forceinline__ float s8_to_f(uint32_t w, int k) { return (float)(int)(int8_t)((w >> (8 * k)) & 0xFFu); }

__device__ __forceinline__ void block_tile_q6k(const uint8_t* bt, const uint4* xfb, const float* offb, float (&tot)[4], int lane, int bsel) {
    const int g = lane >> 2;
    const uint4* q = reinterpret_cast<const uint4*>(bt);
    const uint4 A0 = q[lane], B0 = q[32 + lane], H0 = q[64 + lane];          // row g
    const uint4 A1 = q[96 + lane], B1 = q[128 + lane], H1 = q[160 + lane];   // row g+8
    const uint4 S0 = q[192 + g], S1 = q[200 + g];
    const float4* of4 = reinterpret_cast<const float4*>(offb);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
        const float4 o0 = of4[hf * 2], o1 = of4[hf * 2 + 1];                 // offsets of groups 8hf .. 8hf+7
        const float ofs[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        const uint32_t s0w[2] = {word_of(S0, hf * 2), word_of(S0, hf * 2 + 1)}, s1w[2] = {word_of(S1, hf * 2), word_of(S1, hf * 2 + 1)};
        // per is: the quant words and the shifted qh words of both rows
        uint32_t wA[2][2], wB[2][2], hw[2][2];
#pragma unroll
        for (int is = 0; is < 2; is++) {
            wA[0][is] = word_of(A0, hf * 2 + is); wB[0][is] = word_of(B0, hf * 2 + is); hw[0][is] = word_of(H0, hf * 2 + is);
            wA[1][is] = word_of(A1, hf * 2 + is); wB[1][is] = word_of(B1, hf * 2 + is); hw[1][is] = word_of(H1, hf * 2 + is);
        }
#pragma unroll
        for (int qi = 0; qi < 4; qi++) {
            const uint4 bf = xfb[(hf * 4 + qi) * 12 + bsel];
#pragma unroll
            for (int is = 0; is < 2; is++) {
                const int sgl = 2 * qi + is;   // group inside the half
                uint32_t a[2][2];              // [row half][a0 | a2]
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const uint32_t w = (qi & 1) ? wB[r][is] : wA[r][is], h = hw[r][is];
                    if (qi < 2) {   // low nibbles, qh bits (2qi, 2qi+1) -> bits 4..5
                        const uint32_t hs = qi == 0 ? (h << 4) : (h << 2), hs8 = qi == 0 ? (h >> 4) : (h >> 6);
                        a[r][0] = (w & 0x000F000Fu) | (hs & 0x00300030u);
                        a[r][1] = ((w >> 8) & 0x000F000Fu) | (hs8 & 0x00300030u);
                    } else {        // high nibbles stay at bits 4..7 (n * 2^-20), qh bits (2qi, 2qi+1) -> bits 8..9
                        const uint32_t hs = qi == 2 ? (h << 4) : (h << 2), hs8 = qi == 2 ? (h >> 4) : (h >> 6);
                        a[r][0] = (w & 0x00F000F0u) | (hs & 0x03000300u);
                        a[r][1] = ((w >> 8) & 0x00F000F0u) | (hs8 & 0x03000300u);
                    }
                }
                const float cin[4] = {ofs[sgl], 0.0f, ofs[sgl], 0.0f};
                float c[4];
                mma_f16(c, a[0][0], a[1][0], a[0][1], a[1][1], is ? bf.z : bf.x, is ? bf.w : bf.y, cin);
                const float f0 = s8_to_f(s0w[sgl >> 2], sgl & 3), f1 = s8_to_f(s1w[sgl >> 2], sgl & 3);
                acc[0] = fmaf(f0, c[0], acc[0]); acc[1] = fmaf(f0, c[1], acc[1]);
                acc[2] = fmaf(f1, c[2], acc[2]); acc[3] = fmaf(f1, c[3], acc[3]);
            }
        }
    }
    const float d0 = h2f(*reinterpret_cast<const uint16_t*>(bt + 3328 + 2 * g)), d1 = h2f(*reinterpret_cast<const uint16_t*>(bt + 3344 + 2 * g));
    tot[0] = fmaf(d0, acc[0], tot[0]); tot[1] = fmaf(d0, acc[1], tot[1]);
    tot[2] = fmaf(d1, acc[2], tot[2]); tot[3] = fmaf(d1, acc[3], tot[3]);
}

__device__ __forceinline__ float silu_mul(float gate, float up) {  // silu_generic.go:22-31: sigmoid in f64, float32(g*sig)*u
    const double gv = (double)gate;
    return (float)(gv * (1.0 / (1.0 + exp(-gv)))) * up;
}

template <int TYPE>
__global__ void __launch_bounds__(kMT, 1) gemv_mma_kernel(const uint8_t* __restrict__ wm, int M, int K, const MGeom g, const Prologue p,
                                                          float* __restrict__ y, int pairs, float* __restrict__ gpart, int* __restrict__ tickets) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float red[32];
    constexpr int BT = bt_bytes(TYPE);
    float* xs = reinterpret_cast<float*>(smem);
    float4* xsum = reinterpret_cast<float4*>(smem + g.xsum_off);
    uint4* xf = reinterpret_cast<uint4*>(smem + g.xf_off);
    uint32_t* xm = reinterpret_cast<uint32_t*>(smem + g.xm_off);   // Q6_K: float offsets; lanes with t != 0 read the zero block behind it
    float* part = reinterpret_cast<float*>(smem + g.part_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem + g.ring_off + (size_t)warp * g.stages * BT;
    const uint32_t bar0 = smem_u32(smem + g.bar_off) + warp * kMStagesMax * 8;

    const int i0 = blockIdx.x * g.per_cta, i1 = min(g.total, i0 + g.per_cta);
    const int n_my = (i0 + warp < i1) ? (i1 - i0 - warp + kMW - 1) / kMW : 0;

    if (lane == 0) {
        for (int s = 0; s < g.stages; s++) mbar_init(bar0 + s * 8, 1);
        fence_barrier_init();
    }
    __syncwarp();
    int issued = 0, ist = 0;
    auto issue_next = [&]() {
        if (lane == 0) {
            const uint32_t bar = bar0 + ist * 8;
            mbar_expect_tx(bar, BT);
            bulk_g2s(smem_u32(ring + (size_t)ist * BT), wm + (size_t)(i0 + warp + issued * kMW) * BT, BT, bar);
        }
        issued++;
        if (++ist == g.stages) ist = 0;
    };
    {
        const int pre = min(n_my, g.stages);
        for (int i = 0; i < pre; i++) issue_next();   // weights are constants: stream them before the dependency resolves
    }
    pdl_launch_dependents();
    pdl_wait();

    build_x<kQ4_K, kMT>(p, p.a, K, xs, xsum, red, blockIdx.x == 0, 0u);
    // power-of-two scale that brings max|x| just under 2^14 (fp16 operands, three-term split)
    float mx = 0.0f;
    for (int i4 = threadIdx.x; i4 < (K >> 2); i4 += kMT) {
        const float4 v = *xslot<kQ4_K>(xs, i4);
        mx = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fmaxf(fabsf(v.z), fabsf(v.w)), mx));
    }
    mx = block_max(mx, red);
    int sh = 140 - (int)((__float_as_uint(mx) >> 23) & 0xFFu);
    sh = max(-60, min(110, sh));
    const float s = __uint_as_float((uint32_t)(sh + 127) << 23);
    // 2^48 / s undoes the 2^-24 of the weights and of the Q4_K scales; Q6_K scales are true integers: 2^24 / s
    const float inv = __uint_as_float((uint32_t)((TYPE == kQ4_K ? 48 : 24) - sh + 127) << 23);
    if (TYPE == kQ4_K) {
        build_frags_q4k(xs, xsum, xf, xm, g.nb, s);
    } else {
        build_frags_q6k(xs, reinterpret_cast<uint2*>(xf), reinterpret_cast<float*>(xm), g.nb, s);
        if (threadIdx.x < 16) xm[g.nb * kXmWords + threadIdx.x] = 0u;
    }
    __syncthreads();

    const int nb = g.nb;
    const int tau_first = i0 / nb;
    const int gq = lane >> 2, t = lane & 3;
    const int bsel = lane < 12 ? lane : (lane < 24 ? lane - 12 : lane - 24);   // lanes >= 12 duplicate fragments: their columns are ignored
    const uint32_t msel = (t & 1) ? 0x4342u : 0x4140u;
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    int i = i0 + warp, tau = i / nb, b = i - tau * nb;
    int cur_tau = -1, cur_slot = 0, st = 0;
    uint32_t parity = 0;
    auto flush = [&]() {
        float vlo = t == 0 ? tot[0] + tot[1] : (t == 1 ? tot[0] : 0.0f);
        float vhi = t == 0 ? tot[2] + tot[3] : (t == 1 ? tot[2] : 0.0f);
        vlo += __shfl_xor_sync(0xffffffffu, vlo, 1); vhi += __shfl_xor_sync(0xffffffffu, vhi, 1);
        vlo += __shfl_xor_sync(0xffffffffu, vlo, 2); vhi += __shfl_xor_sync(0xffffffffu, vhi, 2);
        if (t == 0) {
            float* dst = part + (size_t)((cur_tau - tau_first) * g.slots + cur_slot) * 16;
            dst[gq] = vlo;
            dst[gq + 8] = vhi;
        }
    };
    for (int j = 0; j < n_my; j++) {
        if (tau != cur_tau) {
            if (cur_tau >= 0) flush();
            cur_tau = tau;
            cur_slot = i - max(i0, tau * nb);
            tot[0] = tot[1] = tot[2] = tot[3] = 0.0f;
        }
        mbar_wait(bar0 + st * 8, parity);
        if (TYPE == kQ4_K)
            block_tile_q4k(ring + (size_t)st * BT, xf + (size_t)b * 96, xm + b * kXmWords, tot, lane, bsel, msel);
        else
            block_tile_q6k(ring + (size_t)st * BT, xf + (size_t)b * 96, reinterpret_cast<const float*>(xm) + (t == 0 ? b : g.nb) * kXmWords, tot,
                           lane, bsel);
        __syncwarp();
        if (issued < n_my) {   // refill the stage just drained (generic-proxy reads ordered before the async-proxy write)
            fence_proxy_async();
            issue_next();
        }
        i += kMW;
        b += kMW;
        while (b >= nb) { b -= nb; tau++; }
        if (++st == g.stages) { st = 0; parity ^= 1u; }
    }
    if (cur_tau >= 0) flush();
    __syncthreads();

    // ---- per row tile: sum the warps' partials in slot order; finish complete tiles, hand split tiles to the last CTA
    const int n_local = i1 > i0 ? (i1 - 1) / nb - tau_first + 1 : 0;
    for (int base = 0; base < n_local * 16; base += kMT) {
        const int idx = base + threadIdx.x, tl = idx >> 4, row = idx & 15;
        const bool valid = tl < n_local;
        const int tt = tau_first + tl;
        const int lo = max(i0, tt * nb), hi = min(i1, (tt + 1) * nb);
        const int ns = min(hi - lo, kMW);
        float v = 0.0f;
        if (valid)
            for (int k = 0; k < ns; k++) v += part[(size_t)(tl * g.slots + k) * 16 + row];
        v *= inv;
        const bool complete = (lo == tt * nb) && (hi == (tt + 1) * nb);
        bool fin = valid && complete;
        int nparts = 1;
        if (valid && !complete) {
            const int c_first = (tt * nb) / g.per_cta, c_last = ((tt + 1) * nb - 1) / g.per_cta;
            nparts = c_last - c_first + 1;
            gpart[((size_t)tt * kMaxParts + (blockIdx.x - c_first)) * 16 + row] = v;
            __threadfence();
        }
        __syncwarp();
        int old = -1;
        if (valid && !complete && row == 0) old = atomicAdd(&tickets[tt], 1);
        old = __shfl_sync(0xffffffffu, old, lane & 16);
        if (valid && !complete && old == nparts - 1) {   // last CTA of this row tile: fixed summation order over the parts
            __threadfence();
            v = 0.0f;
            for (int pp = 0; pp < nparts; pp++) v += __ldcg(&gpart[((size_t)tt * kMaxParts + pp) * 16 + row]);
            fin = true;
            if (row == 0) tickets[tt] = 0;   // ready for the next launch (stream order)
        }
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
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && v[0]) ? atoi(v) : dflt;
}

bool make_mgeom(int type, int M, int K, MGeom& g) {
    if ((type != kQ4_K && type != kQ6_K) || K % 256 || K <= 0 || M <= 0) return false;
    const int BT = bt_bytes(type);
    g.nb = K / 256;
    g.n_tiles = (M + 15) / 16;
    const long long total = (long long)g.n_tiles * g.nb;
    if (total > (1ll << 30)) return false;
    g.total = (int)total;
    static const int sms = env_int("ZB_MMA_CTAS", ZB_SMS);
    int per = (g.total + sms - 1) / sms;
    const int min_per = (g.nb + kMaxParts - 2) / (kMaxParts - 1);   // a row tile may span at most kMaxParts CTAs
    if (per < min_per) per = min_per;
    g.per_cta = per;
    g.ctas = (g.total + per - 1) / per;
    g.slots = g.nb < kMW ? g.nb : kMW;
    g.max_local = (per + g.nb - 2) / g.nb + 1;
    const int xbytes = ((K / 64) * 68 * 4 + 127) & ~127;
    const int xsum_bytes = ((K / 64) * 16 + 127) & ~127;
    const int xf_bytes = (g.nb * 96 * 16 + 127) & ~127;
    const int xm_bytes = ((g.nb + 1) * kXmWords * 4 + 127) & ~127;   // + one zero block
    const int part_bytes = (g.max_local * g.slots * 64 + 127) & ~127;
    g.xsum_off = xbytes;
    g.xf_off = g.xsum_off + xsum_bytes;
    g.xm_off = g.xf_off + xf_bytes;
    g.part_off = g.xm_off + xm_bytes;
    g.ring_off = g.part_off + part_bytes;
    const int left = kMSmem - 2048 - g.ring_off - kMW * kMStagesMax * 8 - 128;
    if (left < 0) return false;
    g.stages = left / (kMW * BT);
    if (g.stages > kMStagesMax) g.stages = kMStagesMax;
    if (g.stages < 2) return false;
    g.bar_off = (g.ring_off + kMW * g.stages * BT + 15) & ~15;
    g.smem_bytes = g.bar_off + kMW * kMStagesMax * 8;
    return true;
}

}  // namespace

// ===========================================================================
// C ABI (include/zb200.h)
// ===========================================================================
ZB_API int zb_mma_check(int qtype, int rows, int cols) {
    MGeom g{};
    return make_mgeom(qtype, rows, cols, g) ? 0 : (int)cudaErrorInvalidConfiguration;
}

ZB_API int zb_mma_layout(int qtype, int rows, int cols, int64_t* weight_bytes, int64_t* scratch_bytes) {
    MGeom g{};
    if (!make_mgeom(qtype, rows, cols, g)) return cudaErrorInvalidConfiguration;
    if (weight_bytes) *weight_bytes = (int64_t)g.total * bt_bytes(qtype);
    if (scratch_bytes) *scratch_bytes = (((int64_t)g.n_tiles * 4 + 127) & ~(int64_t)127) + (int64_t)g.n_tiles * kMaxParts * 64;
    return 0;
}

// Host-side repack of raw GGUF blocks into block-tiles (pure byte moves; rows past `rows` are zero blocks).
ZB_API int zb_mma_repack_host(int qtype, const void* raw, int rows, int cols, void* out) {
    MGeom g{};
    if (!make_mgeom(qtype, rows, cols, g)) return cudaErrorInvalidConfiguration;
    const uint8_t* src = static_cast<const uint8_t*>(raw);
    uint8_t* dst = static_cast<uint8_t*>(out);
    const int BT = bt_bytes(qtype);
    memset(dst, 0, (size_t)g.total * BT);
    if (qtype == zb::kQ6_K) {
        for (int tau = 0; tau < g.n_tiles; tau++)
            for (int b = 0; b < g.nb; b++) {
                uint8_t* bt = dst + ((size_t)tau * g.nb + b) * BT;
                for (int h = 0; h < 2; h++)
                    for (int gq = 0; gq < 8; gq++) {
                        const int row = tau * 16 + gq + 8 * h;
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
        return 0;
    }
    for (int tau = 0; tau < g.n_tiles; tau++)
        for (int b = 0; b < g.nb; b++) {
            uint8_t* bt = dst + ((size_t)tau * g.nb + b) * BT;
            for (int h = 0; h < 2; h++)
                for (int gq = 0; gq < 8; gq++) {
                    const int row = tau * 16 + gq + 8 * h;
                    if (row >= rows) continue;
                    const uint8_t* blk = src + ((size_t)row * g.nb + b) * 144;
                    memcpy(bt + 2048 + (h * 8 + gq) * 16, blk, 16);   // d, dmin, scales[12]
                    for (int pp = 0; pp < 2; pp++)
                        for (int t = 0; t < 4; t++) {
                            uint8_t* o = bt + ((h * 2 + pp) * 32 + gq * 4 + t) * 16;
                            memcpy(o, blk + 16 + 32 * (2 * pp) + 8 * t, 8);
                            memcpy(o + 8, blk + 16 + 32 * (2 * pp + 1) + 8 * t, 8);
                        }
                }
        }
    return 0;
}

ZB_API int zb_gemv_mma_f32(const zb_mma_weight* w, const zb_prologue* p, float* y, void* scratch, int flags, zb_stream_t stream) {
    if (!w || !p || !y || !scratch || !w->data) return cudaErrorInvalidValue;
    if (p->mix_n > 0 || p->n_wait > 0) return cudaErrorInvalidValue;   // MoE combine / fused TP exchange stay on the CUDA-core kernel
    MGeom g{};
    if (!make_mgeom(w->qtype, w->rows, w->cols, g)) return cudaErrorInvalidConfiguration;
    if (w->epilogue == 1 && (w->rows & 1)) return cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ4_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemv_mma_kernel<zb::kQ6_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMSmem - 2048);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    zb::Prologue pr{p->a, p->r, p->w1, p->w2, p->sum_out, nullptr, 0, 0, p->eps, p->swiglu};
    int* tickets = static_cast<int*>(scratch);
    float* gpart = reinterpret_cast<float*>(static_cast<uint8_t*>(scratch) + (((size_t)g.n_tiles * 4 + 127) & ~(size_t)127));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g.ctas, 1, 1);
    cfg.blockDim = dim3(kMT, 1, 1);
    cfg.dynamicSmemBytes = g.smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (flags & 1) ? 1 : 0;
    if (w->qtype == zb::kQ6_K)
        return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<zb::kQ6_K>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y,
                                  w->epilogue == 1 ? 1 : 0, gpart, tickets);
    return cudaLaunchKernelEx(&cfg, gemv_mma_kernel<zb::kQ4_K>, static_cast<const uint8_t*>(w->data), w->rows, w->cols, g, pr, y,
                              w->epilogue == 1 ? 1 : 0, gpart, tickets);
}
