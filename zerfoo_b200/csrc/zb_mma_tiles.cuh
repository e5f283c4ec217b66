// Block-tile arithmetic of the tensor-core batch-1 GEMV (sm_100a), shared by the per-matrix kernel (gemv_mma.cu) and the
// persistent whole-token decode kernel (decode_mega.cu): fragment builders (x -> balanced base-256 digits), one block-tile
// on the tensor pipe per format, and the host-side work split.  See gemv_mma.cu for the design notes and the reference
// semantics replaced (gemv_q4k.cu:68-160, gemv_q5k.cu:68-177, gemv_q6k.cu:45-152, gemm_q4.cu:89-96).
#pragma once
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

__host__ __device__ constexpr int bt_bytes(int type) { return type == kQ4_K ? 2304 : (type == kQ5_K ? 2816 : (type == kQ6_K ? 3360 : (type == kQ4_0 ? 1152 : 0))); }
// weights per row of a block-tile ("unit"): one K-quant super-block, or four Q4_0 blocks
__host__ __device__ constexpr int unit_weights(int type) { return type == kQ4_0 ? 128 : 256; }
__host__ __device__ constexpr int xf_stride(int type) { return type == kQ4_0 ? 48 : 96; }   // uint4 fragments per unit
__host__ __device__ constexpr int xm_stride(int type) { return 16; }    // 32-bit side values per unit
constexpr int kXmWords = 16;             // per super-block: Q4_K 12 half2 min-term fragments, Q6_K 16 f32 offset terms

struct MSel {            // MoE: blockIdx.y = slot k, expert = sel[k]
    const int* sel;
    long long stride;    // bytes between experts
    int a_stride, y_stride;
    long long gpart_stride;   // uint2 elements of partial-sum scratch per slot
};

struct MGeom {
    int nb;            // super-blocks per row
    int n_tiles;       // 16-row tiles
    int total;         // block-tiles
    int per_cta, ctas;
    int per_warp, chunk;   // block-tiles per warp (contiguous run), block-tiles per ring stage
    int stages, slots, max_local;
    int xf_off, xm_off, xinv_off, part_off, ring_off, bar_off, smem_bytes;
};

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

// Power-of-two scale that brings the super-block's max|x| just under 2^14 (fp16 operands), and the factor that undoes
// it together with the 2^-24 of the subnormal weights (and, Q4_K, of the subnormal scales).
__device__ __forceinline__ float frag_scale(float mx, int unscale_exp, float& inv) {
    int sh = 140 - (int)((__float_as_uint(mx) >> 23) & 0xFFu);
    sh = max(-60, min(100, sh));
    inv = __uint_as_float((uint32_t)(unscale_exp - sh + 127) << 23);
    return __uint_as_float((uint32_t)(sh + 127) << 23);
}

// The lane's 8 activations of super-block b -> fp16 B fragments in the k order the nibble pairs come out in.
// One warp per super-block; lane = G*8 + nib*4 + t owns x[256b + 8*lane .. +8) (group G, nibble plane, t).
//   xf[((b*8 + G*2 + nib)*12 + n*4 + t)] (uint4) = {b0,b1 of MMA j=0, b0,b1 of MMA j=1} of split term n
//   xm[b*16 + n*4 + G] (uint32)                   = half2(term_n(sum x of sub-block 2G), term_n(sum x of sub-block 2G+1))
//   xinv[b]                                        = 2^48 / s_b
struct F8 { float v[8]; };
__device__ __noinline__ void frags_q4k(const F8 xx, int b, int lane, uint4* xf, uint32_t* xm, float* xinv) {
    const int t = lane & 3, nib = (lane >> 2) & 1, G = lane >> 3;
    const float (&x)[8] = xx.v;
    float mx = fmaxf(fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))),
                     fmaxf(fmaxf(fabsf(x[4]), fabsf(x[5])), fmaxf(fabsf(x[6]), fabsf(x[7]))));
    float sum = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));   // the sub-block's sum(x): the four t lanes
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    mx = warp_max(mx);
    float inv;
    const float s = frag_scale(mx, 48, inv);
    if (lane == 0) xinv[b] = inv;
    const float sc = nib ? s * 0.0625f : s;   // high nibbles enter the MMA as n * 2^-20: their x carries the 2^-4
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = x[e] * sc;
    float sv = sum * s * 0.015625f;           // sum of 32 x: 2^-6 keeps it inside fp16 range
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

// One Q4_K block-tile (16 rows x 256 weights) on the tensor cores.  tot[0..1] += row g, columns (2t, 2t+1);
// tot[2..3] += row g+8.  Columns 0..2 are the three fp16 terms of x (the other columns repeat them and are ignored).
// Block-tile bytes: [row half h][group pair p][lane][16 B] nibbles (the lane's bytes 8t..8t+7 of groups 2p and 2p+1), then
// [h][g][16 B] = d | dmin | 12 packed scale bytes.  The group loop stays unrolled: a rolled variant (8-byte operand loads,
// ~170-instruction loop) removed the instruction-fetch stalls ncu shows but ran 15-60 % slower -- with four warps per
// scheduler the overlap of the independent HMMA chains inside one warp matters more (profiles/r01_mma_experiments.md).
__device__ __forceinline__ void block_tile_q4k(const uint8_t* bt, const uint4* xfb, const uint32_t* xmb, float invb, float (&tot)[4],
                                               int lane, int bsel, uint32_t msel) {
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
    const float invm = invb * 3.814697265625e-06f;  // the min path carries 2^-30 (2^-24 mins, 2^-6 sums) against 2^-48: 2^-18
    const float da = h2f((uint16_t)(ha.x & 0xFFFFu)) * invb, dma = h2f((uint16_t)(ha.x >> 16)) * invm;
    const float db = h2f((uint16_t)(hb.x & 0xFFFFu)) * invb, dmb = h2f((uint16_t)(hb.x >> 16)) * invm;
    tot[0] = fmaf(-dma, cm[0], fmaf(da, acc[0], tot[0])); tot[1] = fmaf(-dma, cm[1], fmaf(da, acc[1], tot[1]));
    tot[2] = fmaf(-dmb, cm[2], fmaf(db, acc[2], tot[2])); tot[3] = fmaf(-dmb, cm[3], fmaf(db, acc[3], tot[3]));
}

// ---- Q6_K (gemv_q6k.cu:11-25): 16 scale groups of 16 consecutive weights per super-block, one MMA each ----------------
// Block-tile (3360 B = 16 x 210): [row half h][ql run A | ql run B | qh][lane][16 B], then int8 scales [h][g][16], then fp16 d.
// Lane (g, t) owns, per (half hf, is): the ql words A = ql[64hf + 16is + 4t ..], B = ql[64hf + 32 + 16is + 4t ..] and the qh
// word qh[32hf + 16is + 4t ..]: low nibbles + qh bits (0-1 | 2-3) are q1 | q2, high nibbles + bits (4-5 | 6-7) are q3 | q4.
//   xf as uint2[((b*8 + sg/2)*12 + n*4 + t)*2 + (sg&1)] = (b0, b1) of scale group sg, split term n
//   xm as float[b*16 + sg] = -32 * 2^-24 * s * sum(x of group sg): the "- 32" of every weight, fed in as the MMA's C operand
// lane owns x[256b + 4*lane .. +4) (scale group lane/4) and x[256b + 128 + 4*lane .. +4) (scale group 8 + lane/4)
__device__ __noinline__ void frags_q6k(const F8 xx, int b, int lane, uint2* xf2, float* off, float* xinv) {
    const int t = lane & 3;
    const float (&x)[8] = xx.v;
    float mx = fmaxf(fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))),
                     fmaxf(fmaxf(fabsf(x[4]), fabsf(x[5])), fmaxf(fabsf(x[6]), fabsf(x[7]))));
    mx = warp_max(mx);
    float inv;
    const float s = frag_scale(mx, 24, inv);
    if (lane == 0) xinv[b] = inv;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int sg = i * 8 + (lane >> 2);
        const bool hi = ((sg >> 1) & 3) >= 2;            // q3 | q4 enter as n * 2^-20
        const float sc = hi ? s * 0.0625f : s;
        float v[4] = {x[4 * i] * sc, x[4 * i + 1] * sc, x[4 * i + 2] * sc, x[4 * i + 3] * sc};
        float sum = (x[4 * i] + x[4 * i + 1]) + (x[4 * i + 2] + x[4 * i + 3]);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (t == 0) off[b * kXmWords + sg] = -1.9073486328125e-06f * s * sum;   // 32 * 2^-24 = 2^-19
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

__device__ __forceinline__ void block_tile_q6k(const uint8_t* bt, const uint4* xfb, const float* offb, float invb, float (&tot)[4], int lane,
                                               int bsel) {
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
        // per is: the quant words and the qh words of both rows
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
    const float d0 = h2f(*reinterpret_cast<const uint16_t*>(bt + 3328 + 2 * g)) * invb, d1 = h2f(*reinterpret_cast<const uint16_t*>(bt + 3344 + 2 * g)) * invb;
    tot[0] = fmaf(d0, acc[0], tot[0]); tot[1] = fmaf(d0, acc[1], tot[1]);
    tot[2] = fmaf(d1, acc[2], tot[2]); tot[3] = fmaf(d1, acc[3], tot[3]);
}

// ---- Q4_0 (gemm_q4.cu:1-12,89-96; q4dot.go:10-29): 32-weight blocks, fp16 d + 16 nibble bytes, w = (q - 8) * d --------
// Unit = 16 rows x 4 blocks (1152 B): [row half h][lane][16 B] = word t (bytes 4t..4t+3) of blocks 0..3, then [g][16 B] =
// fp16 d of (row g, blocks 0..3 | row g+8, blocks 0..3).  Low nibbles are weights 0..15 of a block, high nibbles 16..31:
// two MMAs per block share the accumulator (one scale), the "- 8" of every weight enters as the MMA's C operand.
//   xf as uint2[((blk*12 + n*4 + t)*2 + plane)] = (b0, b1);  xm as float[blk] = -8 * 2^-24 * s * sum(x of the block)
// lane owns x[256xb + 8*lane .. +8): block lane/4, elements 8q .. 8q+7 of it, q = lane & 3
__device__ __noinline__ void frags_q40(const F8 xx, int xb, int lane, bool valid, uint2* xf2, float* off, float* xinv) {
    const float (&x)[8] = xx.v;
    const int q = lane & 3, blk = xb * 8 + (lane >> 2), plane = q >> 1;
    float mx = fmaxf(fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))),
                     fmaxf(fmaxf(fabsf(x[4]), fabsf(x[5])), fmaxf(fabsf(x[6]), fabsf(x[7]))));
    mx = warp_max(mx);
    float inv;
    const float s = frag_scale(mx, 24, inv);
    if (lane == 0) xinv[xb] = inv;
    float sum = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (valid && q == 0) off[(blk >> 2) * 16 + (blk & 3)] = -4.76837158203125e-07f * s * sum;   // 8 * 2^-24 = 2^-21
    const float sc = plane ? s * 0.0625f : s;   // high nibbles enter the MMA as n * 2^-20
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = x[e] * sc;
#pragma unroll
    for (int n = 0; n < 3; n++) {
        __half h[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            h[e] = __float2half_rn(v[e]);
            v[e] -= __half2float(h[e]);
        }
        if (valid) {   // elements 8q..8q+3 are lane t = 2(q&1) of the plane's MMA, 8q+4..8q+7 lane t+1
            const int t0 = 2 * (q & 1);
            xf2[(size_t)((blk * 12 + n * 4 + t0) * 2 + plane)] = make_uint2(pack_h2(h[0], h[2]), pack_h2(h[1], h[3]));
            xf2[(size_t)((blk * 12 + n * 4 + t0 + 1) * 2 + plane)] = make_uint2(pack_h2(h[4], h[6]), pack_h2(h[5], h[7]));
        }
    }
}

__device__ __forceinline__ void block_tile_q40(const uint8_t* bt, const uint4* xfb, const float* offb, float invb, float (&tot)[4], int lane,
                                               int bsel) {
    const int g = lane >> 2;
    const uint4* q = reinterpret_cast<const uint4*>(bt);
    const uint4 qa = q[lane], qb = q[32 + lane], sd = q[64 + g];
    const float4 o4 = *reinterpret_cast<const float4*>(offb);
    const float ofs[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
    for (int bi = 0; bi < 4; bi++) {
        const uint32_t wa = word_of(qa, bi), wb = word_of(qb, bi), sa = wa >> 8, sb = wb >> 8;
        const uint4 bf = xfb[bi * 12 + bsel];
        const float cin[4] = {ofs[bi], 0.0f, ofs[bi], 0.0f};
        float c[4];
        mma_f16(c, wa & 0x000F000Fu, wb & 0x000F000Fu, sa & 0x000F000Fu, sb & 0x000F000Fu, bf.x, bf.y, cin);
        mma_f16(c, wa & 0x00F000F0u, wb & 0x00F000F0u, sa & 0x00F000F0u, sb & 0x00F000F0u, bf.z, bf.w, c);
        const float2 da = h2x2_to_f2(bi < 2 ? sd.x : sd.y), db = h2x2_to_f2(bi < 2 ? sd.z : sd.w);
        const float d0 = ((bi & 1) ? da.y : da.x) * invb, d1 = ((bi & 1) ? db.y : db.x) * invb;
        tot[0] = fmaf(d0, c[0], tot[0]); tot[1] = fmaf(d0, c[1], tot[1]);
        tot[2] = fmaf(d1, c[2], tot[2]); tot[3] = fmaf(d1, c[3], tot[3]);
    }
}

// ---- integer tensor-core variant (IMMA m16n8k32, u8 x s8 -> s32) ----------------------------------------------------
// Measured on B200: HMMA.16816 and IMMA.16832 both issue at 0.5 per clock per SM, so the int8 shape does twice the k per
// instruction; a masked quant word (w & 0x0F0F0F0F) IS four u8 operands (one LOP3 per four weights instead of per two);
// and the arithmetic is exact: x enters as a 32-bit fixed-point number per super-block, split into four balanced base-256
// digits in four of the eight B columns, all products and sums are integers (|sum| < 2^25), the 6-bit scales are applied
// with IMAD, and only the final per-super-block conversion is rounded (f32).  No alignment truncation as on the f16 path.
__device__ __forceinline__ void mma_i8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1,
                                       const int (&c)[4]) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
}

// power-of-two scale that brings max|x| of the super-block under 2^30, and its inverse
__device__ __forceinline__ float fixed_scale(float mx, float& inv) {
    int sh = 156 - (int)((__float_as_uint(mx) >> 23) & 0xFFu);
    sh = max(-60, min(120, sh));
    inv = __uint_as_float((uint32_t)(127 - sh) << 23);
    return __uint_as_float((uint32_t)(sh + 127) << 23);
}
// Balanced base-256 digits without a loop: v = sum_j d_j 256^j with d_j in [-128, 127]  <=>  v + 0x80808080 = sum_j (d_j + 128) 256^j
// with every (d_j + 128) a plain byte, so the four s8 digits of v are the bytes of (v + 0x80808080) ^ 0x80808080
// (byte 0 = least significant digit).  |v| <= 2^30 keeps the sum inside 32 bits.
__device__ __forceinline__ uint32_t digit_bytes(float xs) { return ((uint32_t)__float2int_rn(xs) + 0x80808080u) ^ 0x80808080u; }
// 4 x 4 byte transpose: o[j] = (byte 3-j of w0, of w1, of w2, of w3) = digit j (most significant first) of four consecutive elements
__device__ __forceinline__ void digits_of4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t (&o)[4]) {
    const uint32_t t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w0, w1, 0x7362);
    const uint32_t t2 = __byte_perm(w2, w3, 0x5140), t3 = __byte_perm(w2, w3, 0x7362);
    o[3] = __byte_perm(t0, t2, 0x5410); o[2] = __byte_perm(t0, t2, 0x7632);
    o[1] = __byte_perm(t1, t3, 0x5410); o[0] = __byte_perm(t1, t3, 0x7632);
}

// lane = G*8 + nib*4 + t owns x[256b + 8*lane .. +8).
//   xf as uint2[(((b*96/1) ... see below)]: uint2 index ((b*96 + G*16 + j*4 + t)*2 + nib) = (b0, b1) digit j of sub-block 2G+nib
//   xm[b*16 + j*4 + t'] (t' < 2) = digit j of sum(x) of sub-blocks 4t'..4t'+3, one per byte; words with t' >= 2 are zero
//   xinv[b] = 2^-sh
__device__ __forceinline__ void frags_q4k_i8(const F8 xx, int b, int lane, uint4* xf, uint32_t* xm, float* xinv) {
    const float (&x)[8] = xx.v;
    const int t = lane & 3, nib = (lane >> 2) & 1, G = lane >> 3;
    float mx = fmaxf(fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))),
                     fmaxf(fmaxf(fabsf(x[4]), fabsf(x[5])), fmaxf(fabsf(x[6]), fabsf(x[7]))));
    float sum = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    mx = warp_max(mx);
    float inv;
    const float s = fixed_scale(mx, inv);
    if (lane == 0) xinv[b] = inv;
    uint32_t lo[4], hi[4];
    digits_of4(digit_bytes(x[0] * s), digit_bytes(x[1] * s), digit_bytes(x[2] * s), digit_bytes(x[3] * s), lo);
    digits_of4(digit_bytes(x[4] * s), digit_bytes(x[5] * s), digit_bytes(x[6] * s), digit_bytes(x[7] * s), hi);
    uint2* xf2 = reinterpret_cast<uint2*>(xf);
#pragma unroll
    for (int j = 0; j < 4; j++) xf2[(size_t)((b * 96 + G * 16 + j * 4 + t) * 2 + nib)] = make_uint2(lo[j], hi[j]);
    // sum(x) of the sub-block (|sum| <= 32 max|x|: 2^-6 keeps it inside 32 bits): digit words gathered by sub-block quads
    const uint32_t mine = __byte_perm(digit_bytes(sum * s * 0.015625f), 0u, 0x0123);   // bytes = digits 0..3 (most significant first)
    // lane (j, t') needs byte j of the words of sub-blocks 4t' .. 4t'+3, i.e. of lanes 16t' + {0, 4, 8, 12}
    const int jj = (lane >> 2) & 3, tp = lane & 1;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = __shfl_sync(0xffffffffu, mine, 16 * tp + 4 * i);
    if (lane < 16) {
        uint32_t o = 0u;
        if ((lane & 3) < 2) {
#pragma unroll
            for (int i = 0; i < 4; i++) o |= ((w[i] >> (8 * jj)) & 0xFFu) << (8 * i);
        }
        xm[b * kXmWords + lane] = o;
    }
}

// One Q4_K block-tile with integer MMAs.  Per group G: the low nibbles of the lane's two words are a0 | a2 of sub-block 2G,
// the high nibbles of sub-block 2G+1; rows g (a0, a2) and g+8 (a1, a3).  tot[0..1] += row g, digit columns (2t, 2t+1).
// Q5 = true: Q5_K (gemv_q5k.cu:15-23), the same tile followed by [h][lane][8 B] = the lane's qh bytes 8t..8t+7; bit 2G of a qh
// byte is the fifth bit of the group's low-nibble weight, bit 2G+1 of its high-nibble weight.
template <bool Q5>
__device__ __forceinline__ void block_tile_q4k_i8(const uint8_t* bt, const uint4* xfb, const uint32_t* xmb, float invb, float (&tot)[4],
                                                  int lane, float wlo, float whi) {
    const int g = lane >> 2, t = lane & 3, bsel = lane & 15;
    const uint4* q = reinterpret_cast<const uint4*>(bt);
    uint4 qa[2], qb[2];
    qa[0] = q[lane]; qa[1] = q[32 + lane];
    qb[0] = q[64 + lane]; qb[1] = q[96 + lane];
    const uint4 ha = q[128 + g], hb = q[136 + g];
    uint32_t sca[2], scb[2], mna[2], mnb[2];
    sca[0] = ha.y & 0x3F3F3F3Fu; sca[1] = (ha.w & 0x0F0F0F0Fu) | ((ha.y >> 2) & 0x30303030u);
    scb[0] = hb.y & 0x3F3F3F3Fu; scb[1] = (hb.w & 0x0F0F0F0Fu) | ((hb.y >> 2) & 0x30303030u);
    mna[0] = ha.z & 0x3F3F3F3Fu; mna[1] = ((ha.w >> 4) & 0x0F0F0F0Fu) | ((ha.z >> 2) & 0x30303030u);
    mnb[0] = hb.z & 0x3F3F3F3Fu; mnb[1] = ((hb.w >> 4) & 0x0F0F0F0Fu) | ((hb.z >> 2) & 0x30303030u);
    const int zero[4] = {0, 0, 0, 0};
    int acc[4] = {0, 0, 0, 0};
    uint2 qha = make_uint2(0u, 0u), qhb = make_uint2(0u, 0u);
    if (Q5) {
        const uint2* qh = reinterpret_cast<const uint2*>(bt + 2304);
        qha = qh[lane];
        qhb = qh[32 + lane];
    }
#pragma unroll
    for (int G = 0; G < 4; G++) {
        const uint32_t wa0 = (G & 1) ? qa[G >> 1].z : qa[G >> 1].x, wa1 = (G & 1) ? qa[G >> 1].w : qa[G >> 1].y;
        const uint32_t wb0 = (G & 1) ? qb[G >> 1].z : qb[G >> 1].x, wb1 = (G & 1) ? qb[G >> 1].w : qb[G >> 1].y;
        const uint4 bf = xfb[G * 16 + bsel];
        uint32_t l0 = wa0 & 0x0F0F0F0Fu, l1 = wb0 & 0x0F0F0F0Fu, l2 = wa1 & 0x0F0F0F0Fu, l3 = wb1 & 0x0F0F0F0Fu;
        uint32_t h0 = (wa0 >> 4) & 0x0F0F0F0Fu, h1 = (wb0 >> 4) & 0x0F0F0F0Fu, h2 = (wa1 >> 4) & 0x0F0F0F0Fu, h3 = (wb1 >> 4) & 0x0F0F0F0Fu;
        if (Q5) {   // fifth bits: (qh >> 2G) bit 0 -> low plane, bit 1 -> high plane, moved to bit 4 of every byte
            const uint32_t a0 = qha.x >> (2 * G), a1 = qha.y >> (2 * G), b0 = qhb.x >> (2 * G), b1 = qhb.y >> (2 * G);
            l0 |= (a0 << 4) & 0x10101010u; l1 |= (b0 << 4) & 0x10101010u; l2 |= (a1 << 4) & 0x10101010u; l3 |= (b1 << 4) & 0x10101010u;
            h0 |= (a0 << 3) & 0x10101010u; h1 |= (b0 << 3) & 0x10101010u; h2 |= (a1 << 3) & 0x10101010u; h3 |= (b1 << 3) & 0x10101010u;
        }
        int cl[4], ch[4];
        mma_i8(cl, l0, l1, l2, l3, bf.x, bf.y, zero);
        mma_i8(ch, h0, h1, h2, h3, bf.z, bf.w, zero);
        const int k0 = (G & 1) * 2;   // bytes (k0, k0+1) of the scale word = sub-blocks (2G, 2G+1)
        const int sal = (int)__byte_perm(sca[G >> 1], 0u, 0x4440u + k0), sah = (int)__byte_perm(sca[G >> 1], 0u, 0x4441u + k0);
        const int sbl = (int)__byte_perm(scb[G >> 1], 0u, 0x4440u + k0), sbh = (int)__byte_perm(scb[G >> 1], 0u, 0x4441u + k0);
        acc[0] += sal * cl[0] + sah * ch[0]; acc[1] += sal * cl[1] + sah * ch[1];
        acc[2] += sbl * cl[2] + sbh * ch[2]; acc[3] += sbl * cl[3] + sbh * ch[3];
    }
    // min term: A[row][k = sub-block] = m (u8, k 8..31 zero), B[k][j] = digit j of sum(x of sub-block k) * 2^-6
    int cm[4];
    mma_i8(cm, t == 0 ? mna[0] : (t == 1 ? mna[1] : 0u), t == 0 ? mnb[0] : (t == 1 ? mnb[1] : 0u), 0u, 0u, xmb[bsel], 0u, zero);
    const float da = h2f((uint16_t)(ha.x & 0xFFFFu)) * invb, dma = h2f((uint16_t)(ha.x >> 16)) * invb * 64.0f;
    const float db = h2f((uint16_t)(hb.x & 0xFFFFu)) * invb, dmb = h2f((uint16_t)(hb.x >> 16)) * invb * 64.0f;
    // digit column weights: wlo = 2^(8(3-2t)), whi = 2^(8(2-2t)) for t < 2, zero for the unused columns 4..7
    tot[0] = fmaf(wlo, da * (float)acc[0] - dma * (float)cm[0], tot[0]); tot[1] = fmaf(whi, da * (float)acc[1] - dma * (float)cm[1], tot[1]);
    tot[2] = fmaf(wlo, db * (float)acc[2] - dmb * (float)cm[2], tot[2]); tot[3] = fmaf(whi, db * (float)acc[3] - dmb * (float)cm[3], tot[3]);
}

// ---- Q6_K, integer path.  One IMMA per (half hf, q-plane qi) covers BOTH 16-weight scale groups (is = 0, 1) of the plane:
// group is=0 sits in k 0..15 and meets its x digits in B columns 0..3 (zeros in k 16..31), group is=1 in k 16..31 and
// columns 4..7 -- all eight columns carry useful sums.  u8 operand = (ql nibble) | (qh bits << 4), the "- 32" enters as the
// integer C operand, int8 scales are applied with IMAD.
//   words of super-block b (xf + b*96 as uint32): [m*32 + (is*4 + j)*4 + t] = digit j of x[16*sg + 4t .. +4), m = hf*4 + qi,
//   sg = 8hf + 2qi + is;  [256 + m*8 + is*4 + j] = -32 * sum over the group of digit j
__device__ __forceinline__ void frags_q6k_i8(const F8 xx, int b, int lane, uint4* xf, float* xinv) {
    const float (&x)[8] = xx.v;
    const int t = lane & 3;
    float mx = fmaxf(fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))),
                     fmaxf(fmaxf(fabsf(x[4]), fabsf(x[5])), fmaxf(fabsf(x[6]), fabsf(x[7]))));
    mx = warp_max(mx);
    float inv;
    const float s = fixed_scale(mx, inv);
    if (lane == 0) xinv[b] = inv;
    uint32_t* xw = reinterpret_cast<uint32_t*>(xf + (size_t)b * 96);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int sg = i * 8 + (lane >> 2), hf = i, qi = (sg >> 1) & 3, is = sg & 1, m = hf * 4 + qi;
        uint32_t dw[4];
        digits_of4(digit_bytes(x[4 * i] * s), digit_bytes(x[4 * i + 1] * s), digit_bytes(x[4 * i + 2] * s), digit_bytes(x[4 * i + 3] * s), dw);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            xw[m * 32 + (is * 4 + j) * 4 + t] = dw[j];
            int sum = __dp4a((int)dw[j], 0x01010101, 0);   // the four s8 digits of this lane
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (t == 0) reinterpret_cast<int*>(xw)[256 + m * 8 + is * 4 + j] = -32 * sum;
        }
    }
}

template <int QI>
__device__ __forceinline__ uint32_t q6_u8(uint32_t w, uint32_t h) {
    if (QI == 0) return (w & 0x0F0F0F0Fu) | ((h << 4) & 0x30303030u);
    if (QI == 1) return (w & 0x0F0F0F0Fu) | ((h << 2) & 0x30303030u);
    if (QI == 2) return ((w >> 4) & 0x0F0F0F0Fu) | (h & 0x30303030u);
    return ((w >> 4) & 0x0F0F0F0Fu) | ((h >> 2) & 0x30303030u);
}

template <int HF, int QI>
__device__ __forceinline__ void q6k_i8_step(const uint4& A0, const uint4& B0, const uint4& H0, const uint4& A1, const uint4& B1, const uint4& H1,
                                            const uint4& S0, const uint4& S1, const uint32_t* xw, int lane, uint32_t selA, uint32_t selB,
                                            uint32_t m0, int (&acc)[4]) {
    constexpr int m = HF * 4 + QI;
    const uint32_t w00 = word_of((QI & 1) ? B0 : A0, HF * 2), w01 = word_of((QI & 1) ? B0 : A0, HF * 2 + 1);   // row g: is 0, 1
    const uint32_t w10 = word_of((QI & 1) ? B1 : A1, HF * 2), w11 = word_of((QI & 1) ? B1 : A1, HF * 2 + 1);   // row g+8
    const uint32_t h00 = word_of(H0, HF * 2), h01 = word_of(H0, HF * 2 + 1), h10 = word_of(H1, HF * 2), h11 = word_of(H1, HF * 2 + 1);
    const uint32_t bw = xw[m * 32 + lane];
    const int2 oc = *reinterpret_cast<const int2*>(xw + 256 + m * 8 + 2 * (lane & 3));
    const int cin[4] = {oc.x, oc.y, oc.x, oc.y};
    int c[4];
    mma_i8(c, q6_u8<QI>(w00, h00), q6_u8<QI>(w10, h10), q6_u8<QI>(w01, h01), q6_u8<QI>(w11, h11), bw & m0, bw & ~m0, cin);
    const uint32_t shl = (QI & 1) ? selB : selA;   // 24 - 8 * (byte index of the lane's scale): shift it to the top, arithmetic shift back
    const int s0 = (int)(word_of(S0, 2 * HF + (QI >> 1)) << shl) >> 24, s1 = (int)(word_of(S1, 2 * HF + (QI >> 1)) << shl) >> 24;
    acc[0] += s0 * c[0]; acc[1] += s0 * c[1];
    acc[2] += s1 * c[2]; acc[3] += s1 * c[3];
}

__device__ __forceinline__ void block_tile_q6k_i8(const uint8_t* bt, const uint4* xfb, float invb, float (&tot)[4], int lane, uint32_t selA,
                                                  uint32_t selB, float wlo, float whi) {
    const int g = lane >> 2;
    const uint4* q = reinterpret_cast<const uint4*>(bt);
    const uint4 A0 = q[lane], B0 = q[32 + lane], H0 = q[64 + lane];          // row g
    const uint4 A1 = q[96 + lane], B1 = q[128 + lane], H1 = q[160 + lane];   // row g+8
    const uint4 S0 = q[192 + g], S1 = q[200 + g];
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(xfb);
    const uint32_t m0 = lane < 16 ? 0xFFFFFFFFu : 0u;
    int acc[4] = {0, 0, 0, 0};
    q6k_i8_step<0, 0>(A0, B0, H0, A1, B1, H1, S0, S1, xw, lane, selA, selB, m0, acc);
    q6k_i8_step<0, 1>(A0, B0, H0, A1, B1, H1, S0, S1, xw, lane, selA, selB, m0, acc);
    q6k_i8_step<0, 2>(A0, B0, H0, A1, B1, H1, S0, S1, xw, lane, selA, selB, m0, acc);
    q6k_i8_step<0, 3>(A0, B0, H0, A1, B1, H1, S0, S1, xw, lane, selA, selB, m0, acc);
    q6k_i8_step<1, 0>(A0, B0, H0, A1, B1, H1, S0, S1, xw, lane, selA, selB, m0, acc);
    q6k_i8_step<1, 1>(A0, B0, H0, A1, B1, H1, S0, S1, xw, lane, selA, selB, m0, acc);
    q6k_i8_step<1, 2>(A0, B0, H0, A1, B1, H1, S0, S1, xw, lane, selA, selB, m0, acc);
    q6k_i8_step<1, 3>(A0, B0, H0, A1, B1, H1, S0, S1, xw, lane, selA, selB, m0, acc);
    const float d0 = h2f(*reinterpret_cast<const uint16_t*>(bt + 3328 + 2 * g)) * invb, d1 = h2f(*reinterpret_cast<const uint16_t*>(bt + 3344 + 2 * g)) * invb;
    tot[0] = fmaf(wlo * d0, (float)acc[0], tot[0]); tot[1] = fmaf(whi * d0, (float)acc[1], tot[1]);
    tot[2] = fmaf(wlo * d1, (float)acc[2], tot[2]); tot[3] = fmaf(whi * d1, (float)acc[3], tot[3]);
}

// ---- Q4_0, integer path: one IMMA per 32-weight block (low nibbles = k 0..15 = weights 0..15, high nibbles = k 16..31 =
// weights 16..31), the "- 8" as the integer C operand, the fp16 block scale applied after the exact integer dot product.
//   words (xf as uint32, 48 per block... see strides): [blk*12*4 ...] kept simple: uint2 index (blk*16 + j*4 + t) = (b0, b1) digit j
//   xm as int[blk*4 + j] = -8 * sum over the block of digit j
__device__ __forceinline__ void frags_q40_i8(const F8 xx, int xb, int lane, bool valid, uint2* xf2, int* off, float* xinv) {
    const float (&x)[8] = xx.v;
    const int q = lane & 3, blk = xb * 8 + (lane >> 2);
    float mx = fmaxf(fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))),
                     fmaxf(fmaxf(fabsf(x[4]), fabsf(x[5])), fmaxf(fabsf(x[6]), fabsf(x[7]))));
    mx = warp_max(mx);
    float inv;
    const float s = fixed_scale(mx, inv);
    if (lane == 0) xinv[xb] = inv;
    uint32_t lo[4], hi[4];
    digits_of4(digit_bytes(x[0] * s), digit_bytes(x[1] * s), digit_bytes(x[2] * s), digit_bytes(x[3] * s), lo);
    digits_of4(digit_bytes(x[4] * s), digit_bytes(x[5] * s), digit_bytes(x[6] * s), digit_bytes(x[7] * s), hi);
    // the lane's elements 8q..8q+7 of the block: q = 0, 1 are weights 0..15 (b0 of lanes t = 2q, 2q+1), q = 2, 3 weights 16..31 (b1)
    const int t0 = 2 * (q & 1), hp = q >> 1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int sum = __dp4a((int)lo[j], 0x01010101, __dp4a((int)hi[j], 0x01010101, 0));
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (valid) {
            uint32_t* w = reinterpret_cast<uint32_t*>(xf2 + (size_t)(blk * 16 + j * 4 + t0));
            w[hp] = lo[j];
            w[2 + hp] = hi[j];
            if (q == 0) off[blk * 4 + j] = -8 * sum;
        }
    }
}

__device__ __forceinline__ void block_tile_q40_i8(const uint8_t* bt, const uint2* xfu, const int* offu, float invb, float (&tot)[4], int lane,
                                                  float wlo, float whi) {
    const int g = lane >> 2, t = lane & 3, bsel = lane & 15;
    const uint4* q = reinterpret_cast<const uint4*>(bt);
    const uint4 qa = q[lane], qb = q[32 + lane], sd = q[64 + g];
    const float il = invb * wlo, ih = invb * whi;
#pragma unroll
    for (int bi = 0; bi < 4; bi++) {
        const uint32_t wa = word_of(qa, bi), wb = word_of(qb, bi);
        const uint2 bf = xfu[bi * 16 + bsel];
        const int2 oc = *reinterpret_cast<const int2*>(offu + bi * 4 + 2 * (t & 1));
        const int cin[4] = {oc.x, oc.y, oc.x, oc.y};
        int c[4];
        mma_i8(c, wa & 0x0F0F0F0Fu, wb & 0x0F0F0F0Fu, (wa >> 4) & 0x0F0F0F0Fu, (wb >> 4) & 0x0F0F0F0Fu, bf.x, bf.y, cin);
        const float2 da = h2x2_to_f2(bi < 2 ? sd.x : sd.y), db = h2x2_to_f2(bi < 2 ? sd.z : sd.w);
        const float d0 = (bi & 1) ? da.y : da.x, d1 = (bi & 1) ? db.y : db.x;
        tot[0] = fmaf(d0 * il, (float)c[0], tot[0]); tot[1] = fmaf(d0 * ih, (float)c[1], tot[1]);
        tot[2] = fmaf(d1 * il, (float)c[2], tot[2]); tot[3] = fmaf(d1 * ih, (float)c[3], tot[3]);
    }
}

__device__ __noinline__ float silu_mul(float gate, float up) {  // silu_generic.go:22-31: sigmoid in f64, float32(g*sig)*u
    const double gv = (double)gate;
    return (float)(gv * (1.0 / (1.0 + exp(-gv)))) * up;
}

constexpr int kMaxOwn = 4;   // super-blocks of x per warp: K <= kMW * kMaxOwn * 256

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && v[0]) ? atoi(v) : dflt;
}

bool make_mgeom(int type, int M, int K, MGeom& g, int max_ctas = ZB_SMS, bool whole_tiles = false) {
    if ((type != kQ4_K && type != kQ5_K && type != kQ6_K && type != kQ4_0) || K <= 0 || K % unit_weights(type) || M <= 0) return false;
    const int BT = bt_bytes(type);
    g.nb = K / unit_weights(type);
    g.n_tiles = (M + 15) / 16;
    const long long total = (long long)g.n_tiles * g.nb;
    if (total > (1ll << 30)) return false;
    g.total = (int)total;
    static const int sms_env = env_int("ZB_MMA_CTAS", ZB_SMS);
    const int sms = max_ctas < sms_env ? max_ctas : sms_env;
    int per = (g.total + sms - 1) / sms;
    const int min_per = (g.nb + kMaxParts - 2) / (kMaxParts - 1);   // a row tile may span at most kMaxParts CTAs
    if (per < min_per) per = min_per;
    if (whole_tiles) per = ((g.n_tiles + sms - 1) / sms) * g.nb;   // CTAs own whole row tiles: every sum is finished where it is computed
    g.per_cta = per;
    g.ctas = (g.total + per - 1) / per;
    if (K > kMW * kMaxOwn * 256) return false;
    g.per_warp = (per + kMW - 1) / kMW;
    g.slots = (g.nb + g.per_warp - 1) / g.per_warp + 1;   // warps whose runs can touch one row tile
    if (g.slots > kMW) g.slots = kMW;
    g.max_local = (per + g.nb - 2) / g.nb + 1;
    const int xf_bytes = (g.nb * xf_stride(type) * 16 + 127) & ~127;
    const int xm_bytes = (g.nb * xm_stride(type) * 4 + 64 + 127) & ~127;   // + one zero block
    const int xinv_bytes = (((K + 255) / 256) * 4 + 127) & ~127;
    const int part_bytes = (g.max_local * g.slots * 64 + 127) & ~127;
    g.xf_off = 0;
    g.xm_off = g.xf_off + xf_bytes;
    g.xinv_off = g.xm_off + xm_bytes;
    g.part_off = g.xinv_off + xinv_bytes;
    g.ring_off = g.part_off + part_bytes;
    const int left = kMSmem - 2048 - g.ring_off - kMW * kMStagesMax * 8 - 128;
    if (left < 0) return false;
    // ring: stages of `chunk` consecutive block-tiles; two tiles per stage halve the per-tile wait/refill overhead once
    // a warp has enough of them and two such stages fit
    static const int force_chunk = env_int("ZB_MMA_CHUNK", 0);
    g.chunk = 1;
    for (int c = 4; c > 1; c >>= 1)
        if (c * BT <= 4608 && g.per_warp >= 2 * c && left / (kMW * c * BT) >= 2) { g.chunk = c; break; }
    if (force_chunk > 0) g.chunk = force_chunk;
    g.stages = left / (kMW * g.chunk * BT);
    if (g.stages > kMStagesMax) g.stages = kMStagesMax;
    if (g.stages < 2) return false;
    g.bar_off = (g.ring_off + kMW * g.stages * g.chunk * BT + 15) & ~15;
    g.smem_bytes = g.bar_off + kMW * kMStagesMax * 8;
    return true;
}

}  // namespace
