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
#include "zb_stream.cuh"
#include "zb200.h"

namespace {

using namespace zb;

constexpr int kSWarps = 8;
constexpr int kSThreads = kSWarps * 32;
constexpr int kMaxStages = 4;
constexpr int kSmemBudget = 110 * 1024;  // per CTA, so that two stage kernels co-reside on one SM (PDL overlap)
constexpr int kRingBudget = 96 * 1024;

struct SGeom {
    int C;                   // 32-weight chunks per row
    int lpr;                 // lanes sharing a row (4..32)
    int rows_pass;           // rows per tile = (32/lpr) * R
    int cpl;                 // chunks per lane per slab
    int slab_chunks, n_slabs;
    int contig;              // slab == whole row: a tile is one contiguous span of the matrix
    int row_main, row_aux;   // bytes per matrix row
    int slab_main;           // main bytes of one full row-slab
    int slab_main_cap, slab_aux_cap;  // per-row slot sizes inside a stage (slab mode)
    int stage_main, stage_bytes;      // aux region starts at stage_main
    int stages, n_tiles;
    int xsum_off, ring_off, bar_off, smem_bytes;
};

struct Indirect {            // MoE: blockIdx.y = slot k, expert = sel[k]
    const int* sel;
    long long main_stride, aux_stride;
    int a_stride, y_stride;
};

// position of element k of x inside shared memory: chunk-major in the order the
// format's dot product consumes it, 16-B groups XOR-swizzled by chunk.
template <int TYPE>
__device__ __forceinline__ int xpos(int k) {
    int c, p;
    if (TYPE == kQ4_K || TYPE == kQ5_K) {
        int b = k >> 8, e = k & 255;
        c = (b << 3) + ((e >> 6) << 1) + ((e >> 4) & 1);
        p = (((e >> 5) & 1) << 4) + (e & 15);
    } else if (TYPE == kQ6_K) {
        int b = k >> 8, e = k & 255, l = e & 31;
        c = (b << 3) + ((e >> 7) << 2) + (l >> 3);
        p = (((e >> 5) & 3) << 3) + (l & 7);
    } else {
        c = k >> 5;
        p = k & 31;
    }
    return (c << 5) + ((((p >> 2) ^ (c & 7))) << 2) + (p & 3);
}

__device__ __forceinline__ float inv_rms(float sumsq, int D, float eps) {
    return (float)(1.0 / sqrt((double)(sumsq / (float)D + eps)));  // rmsnorm_generic.go:17 (f64 sqrt, one rounding)
}

template <int TYPE>
__device__ void build_x(const Prologue& p, const float* __restrict__ a, int K, float* xs, float2* xsum, float* red, bool lead) {
    const int tid = threadIdx.x;
    if (p.swiglu) {  // silu_generic.go:22-31: sigmoid in f64, float32(g*sig)*u
        for (int i = tid; i < K; i += kSThreads) {
            double gv = (double)a[i];
            double sig = 1.0 / (1.0 + exp(-gv));
            xs[xpos<TYPE>(i)] = (float)(gv * sig) * a[K + i];
        }
    } else {
        float ss = 0.0f;
        for (int i = tid; i < K; i += kSThreads) {
            float v;
            if (p.mix_n > 0) {
                v = 0.0f;
                for (int k = 0; k < p.mix_n; k++) v = v + a[(size_t)k * p.mix_stride + i] * p.mix_w[k];
            } else {
                v = a[i];
            }
            if (!p.w1 && p.r) {
                v = v + p.r[i];
                if (lead && p.sum_out) p.sum_out[i] = v;
            }
            xs[xpos<TYPE>(i)] = v;
            ss = fmaf(v, v, ss);
        }
        if (p.w1) {
            float s1 = inv_rms(block_sum(ss, red), K, p.eps);
            ss = 0.0f;
            for (int i = tid; i < K; i += kSThreads) {
                int at = xpos<TYPE>(i);
                float v = xs[at] * s1 * p.w1[i];
                if (p.r) {
                    v = v + p.r[i];
                    if (lead && p.sum_out) p.sum_out[i] = v;
                }
                xs[at] = v;
                ss = fmaf(v, v, ss);
            }
        }
        if (p.w2) {
            float s2 = inv_rms(block_sum(ss, red), K, p.eps);
            for (int i = tid; i < K; i += kSThreads) {
                int at = xpos<TYPE>(i);
                xs[at] = xs[at] * s2 * p.w2[i];
            }
        }
    }
    __syncthreads();
    if (TYPE == kQ4_K || TYPE == kQ5_K) {  // per-chunk sums of x for the dmin term: (sum of the 16 low-nibble x, sum of the 16 high-nibble x)
        int C = K >> 5;
        for (int c = tid; c < C; c += kSThreads) {
            const float4* xp = reinterpret_cast<const float4*>(xs + (c << 5));
            float sa = 0.0f, sb = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float4 t = xp[j ^ (c & 7)];
                sa += (t.x + t.y) + (t.z + t.w);
                float4 u = xp[(4 + j) ^ (c & 7)];
                sb += (u.x + u.y) + (u.z + u.w);
            }
            xsum[c] = make_float2(sa, sb);
        }
        __syncthreads();
    }
}

// ---- in-register dequantisation ------------------------------------------------
// four masked bytes (each < 128) -> two f32x2 pairs holding 128+b, then minus `bias`
#define ZB_C43 0x43000000u
__device__ __forceinline__ void bytes_to_pairs(uint32_t m, uint64_t nbias, uint64_t& p0, uint64_t& p1) {
    p0 = add2(pack2u(__byte_perm(m, ZB_C43, 0x7044), __byte_perm(m, ZB_C43, 0x7144)), nbias);
    p1 = add2(pack2u(__byte_perm(m, ZB_C43, 0x7244), __byte_perm(m, ZB_C43, 0x7344)), nbias);
}

__device__ __forceinline__ void kq_scale_min2(uint32_t s0, uint32_t s1, uint32_t s2, int j, float& sc, float& mn) {
    uint32_t a, b;  // gemv_q4k.cu:38-56
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

// One 32-weight chunk `cl` of a row-slab held in shared memory against the x chunk in registers.
template <int TYPE>
__device__ __forceinline__ float chunk_dot(const uint8_t* rowm, const uint8_t* rowa, int cl, const uint64_t (&xv)[16], float2 xs) {
    if (TYPE == kQ4_0) {
        uint4 q = *reinterpret_cast<const uint4*>(rowm + cl * 16);
        float d = h2f(*reinterpret_cast<const uint16_t*>(rowa + cl * 2));
        const uint64_t nb = pack2(-136.0f, -136.0f);  // 128 (float trick) + 8 (Q4_0 offset)
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint64_t acc = 0ull;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint64_t a0, a1, b0, b1;
            bytes_to_pairs(w[i] & 0x0F0F0F0Fu, nb, a0, a1);
            bytes_to_pairs((w[i] >> 4) & 0x0F0F0F0Fu, nb, b0, b1);
            acc = fma2(a0, xv[2 * i], acc);
            acc = fma2(a1, xv[2 * i + 1], acc);
            acc = fma2(b0, xv[8 + 2 * i], acc);
            acc = fma2(b1, xv[8 + 2 * i + 1], acc);
        }
        return sum2(acc) * d;
    } else if (TYPE == kQ8_0) {
        const uint4* qp = reinterpret_cast<const uint4*>(rowm + cl * 32);
        float d = h2f(*reinterpret_cast<const uint16_t*>(rowa + cl * 2));
        const uint64_t nb = pack2(-8388736.0f, -8388736.0f);  // 2^23 + 128
        uint64_t acc = 0ull;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint4 q = qp[h];
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t u = w[i] ^ 0x80808080u;  // int8 -> biased uint8
                uint64_t p0 = add2(pack2u(__byte_perm(u, 0x4B000000u, 0x7440), __byte_perm(u, 0x4B000000u, 0x7441)), nb);
                uint64_t p1 = add2(pack2u(__byte_perm(u, 0x4B000000u, 0x7442), __byte_perm(u, 0x4B000000u, 0x7443)), nb);
                acc = fma2(p0, xv[8 * h + 2 * i], acc);
                acc = fma2(p1, xv[8 * h + 2 * i + 1], acc);
            }
        }
        return sum2(acc) * d;
    } else if (TYPE == kQ4_K || TYPE == kQ5_K) {
        constexpr int BB = TYPE == kQ4_K ? 144 : 176;
        const uint8_t* blk = rowm + (cl >> 3) * BB;
        int sub = cl & 7, g = sub >> 1;
        uint4 hdr = *reinterpret_cast<const uint4*>(blk);
        uint4 q = *reinterpret_cast<const uint4*>(blk + 16 + sub * 16);
        float d = h2f((uint16_t)(hdr.x & 0xFFFFu)), dmin = h2f((uint16_t)(hdr.x >> 16));
        // The 8 lanes that share this super-block each decode ONE 6-bit (scale, min) pair -- sub-block `sub`
        // (gemv_q4k.cu:38-56) -- and trade it with the neighbour lane: a chunk needs sub-blocks 2g and 2g+1.
        const int jj = sub & 3;
        uint32_t x0 = __byte_perm(hdr.y, 0u, 0x4440 + jj), x1 = __byte_perm(hdr.z, 0u, 0x4440 + jj), x2 = __byte_perm(hdr.w, 0u, 0x4440 + jj);
        uint32_t scq = sub < 4 ? (x0 & 63u) : ((x2 & 0xFu) | ((x0 >> 6) << 4));
        uint32_t mnq = sub < 4 ? (x1 & 63u) : ((x2 >> 4) | ((x1 >> 6) << 4));
        float my_ds = d * (float)scq, my_dm = dmin * (float)mnq;  // exact products (fp16 x 6-bit)
        float ot_ds = __shfl_xor_sync(0xffffffffu, my_ds, 1), ot_dm = __shfl_xor_sync(0xffffffffu, my_dm, 1);
        const bool odd = sub & 1;
        float ds0 = odd ? ot_ds : my_ds, dm0 = odd ? ot_dm : my_dm, ds1 = odd ? my_ds : ot_ds, dm1 = odd ? my_dm : ot_dm;
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t hb[4] = {0, 0, 0, 0};
        if (TYPE == kQ5_K) {
            uint4 h = *reinterpret_cast<const uint4*>(blk + 144 + (sub & 1) * 16);
            hb[0] = h.x >> (2 * g); hb[1] = h.y >> (2 * g); hb[2] = h.z >> (2 * g); hb[3] = h.w >> (2 * g);
        }
        const uint64_t nb = pack2(-128.0f, -128.0f);
        uint64_t accA = 0ull, accB = 0ull;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t lo = w[i] & 0x0F0F0F0Fu, hi = (w[i] >> 4) & 0x0F0F0F0Fu;
            if (TYPE == kQ5_K) {
                lo |= (hb[i] & 0x01010101u) << 4;
                hi |= (hb[i] & 0x02020202u) << 3;
            }
            uint64_t a0, a1, b0, b1;
            bytes_to_pairs(lo, nb, a0, a1);
            bytes_to_pairs(hi, nb, b0, b1);
            accA = fma2(a0, xv[2 * i], accA);
            accA = fma2(a1, xv[2 * i + 1], accA);
            accB = fma2(b0, xv[8 + 2 * i], accB);
            accB = fma2(b1, xv[8 + 2 * i + 1], accB);
        }
        // sum (d*sc*q - dmin*m) x = d*sc*sum(q x) - dmin*m*sum(x); d*sc and dmin*m are exact in f32
        return ds0 * sum2(accA) - dm0 * xs.x + ds1 * sum2(accB) - dm1 * xs.y;
    } else {  // kQ6_K, split layout: ql[128] qh[64] sc[16] per block, fp16 d in aux
        const uint8_t* blk = rowm + (cl >> 3) * 208;
        int sub = cl & 7, half = sub >> 2, l0 = (sub & 3) * 8;
        uint2 A = *reinterpret_cast<const uint2*>(blk + half * 64 + l0);
        uint2 B = *reinterpret_cast<const uint2*>(blk + half * 64 + 32 + l0);
        uint2 H = *reinterpret_cast<const uint2*>(blk + 128 + half * 32 + l0);
        const int8_t* sc = reinterpret_cast<const int8_t*>(blk + 192) + half * 8 + ((sub & 3) >> 1);
        float d = h2f(*reinterpret_cast<const uint16_t*>(rowa + (cl >> 3) * 2));
        float s1 = d * (float)sc[0], s2 = d * (float)sc[2], s3 = d * (float)sc[4], s4 = d * (float)sc[6];
        const uint64_t nb = pack2(-160.0f, -160.0f);  // 128 + 32
        uint32_t a[2] = {A.x, A.y}, b[2] = {B.x, B.y}, h[2] = {H.x, H.y};
        uint64_t c1 = 0ull, c2 = 0ull, c3 = 0ull, c4 = 0ull;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            uint32_t q1 = (a[i] & 0x0F0F0F0Fu) | ((h[i] << 4) & 0x30303030u);
            uint32_t q2 = (b[i] & 0x0F0F0F0Fu) | ((h[i] << 2) & 0x30303030u);
            uint32_t q3 = ((a[i] >> 4) & 0x0F0F0F0Fu) | (h[i] & 0x30303030u);
            uint32_t q4 = ((b[i] >> 4) & 0x0F0F0F0Fu) | ((h[i] >> 2) & 0x30303030u);
            uint64_t p0, p1;
            bytes_to_pairs(q1, nb, p0, p1);
            c1 = fma2(p0, xv[2 * i], c1);
            c1 = fma2(p1, xv[2 * i + 1], c1);
            bytes_to_pairs(q2, nb, p0, p1);
            c2 = fma2(p0, xv[4 + 2 * i], c2);
            c2 = fma2(p1, xv[4 + 2 * i + 1], c2);
            bytes_to_pairs(q3, nb, p0, p1);
            c3 = fma2(p0, xv[8 + 2 * i], c3);
            c3 = fma2(p1, xv[8 + 2 * i + 1], c3);
            bytes_to_pairs(q4, nb, p0, p1);
            c4 = fma2(p0, xv[12 + 2 * i], c4);
            c4 = fma2(p1, xv[12 + 2 * i + 1], c4);
        }
        return s1 * sum2(c1) + s2 * sum2(c2) + s3 * sum2(c3) + s4 * sum2(c4);
    }
}

// ---- the kernel --------------------------------------------------------------
struct TileCursor {
    int ti, s;  // row-tile ordinal of this warp, slab
};

template <int TYPE, int R>
__global__ void __launch_bounds__(kSThreads, 2) gemv_stream_kernel(StreamW w, const SGeom g, const Prologue p, float* __restrict__ y,
                                                                   const Indirect ind) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float red[32];
    float* xs = reinterpret_cast<float*>(smem);
    float2* xsum = reinterpret_cast<float2*>(smem + g.xsum_off);
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
        if (g.contig) {
            if (lane == 0) {
                uint32_t mb = (uint32_t)nrows * g.row_main;
                uint32_t ab = g.row_aux ? (uint32_t)((nrows * g.row_aux + 16 + 15) & ~15) : 0u;
                mbar_expect_tx(bar, mb + ab);
                bulk_g2s(smem_u32(sm), w.main + (size_t)r0 * g.row_main, mb, bar);
                if (ab) {
                    uintptr_t src = reinterpret_cast<uintptr_t>(w.aux + (size_t)r0 * g.row_aux);
                    bulk_g2s(smem_u32(sm + g.stage_main), reinterpret_cast<const void*>(src & ~(uintptr_t)15), ab, bar);
                }
            }
        } else {
            const int nch = min(g.slab_chunks, g.C - s * g.slab_chunks);
            const uint32_t mb = (uint32_t)stream_main_bytes(w.type, nch);
            const int sa = stream_aux_bytes(w.type, nch);
            const uint32_t ab = sa ? (uint32_t)((sa + 16 + 15) & ~15) : 0u;
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)nrows * (mb + ab));
            __syncwarp();
            for (int rl = lane; rl < nrows; rl += 32) {
                const size_t row = (size_t)(r0 + rl);
                bulk_g2s(smem_u32(sm + rl * g.slab_main_cap), w.main + row * g.row_main + (size_t)s * g.slab_main, mb, bar);
                if (ab) {
                    uintptr_t src = reinterpret_cast<uintptr_t>(w.aux + row * g.row_aux + stream_aux_bytes(w.type, s * g.slab_chunks));
                    bulk_g2s(smem_u32(sm + g.stage_main + rl * g.slab_aux_cap), reinterpret_cast<const void*>(src & ~(uintptr_t)15), ab, bar);
                }
            }
        }
    };

    TileCursor ic{0, 0};  // issue cursor
    int issued = 0;
    auto issue_next = [&]() {
        issue(ic.ti, ic.s, issued % g.stages);
        issued++;
        if (++ic.s == g.n_slabs) { ic.s = 0; ic.ti++; }
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
        w.main += (size_t)e * ind.main_stride;
        if (w.aux) w.aux += (size_t)e * ind.aux_stride;
        a += (size_t)blockIdx.y * ind.a_stride;
        y += (size_t)blockIdx.y * ind.y_stride;
        const int pre = min(nq, g.stages);
        for (int i = 0; i < pre; i++) issue_next();
    }
    build_x<TYPE>(p, a, w.K, xs, xsum, red, blockIdx.x == 0 && blockIdx.y == 0);

    const int sr = lane / g.lpr, lr = lane % g.lpr;
    float acc[R];
    int ti = 0, s = 0;
    for (int q = 0; q < nq; q++) {
        const int st = q % g.stages;
        const uint32_t parity = (uint32_t)(q / g.stages) & 1u;
        const int r0 = (gw + ti * total_w) * g.rows_pass;
        if (s == 0) {
#pragma unroll
            for (int j = 0; j < R; j++) acc[j] = 0.0f;
        }
        mbar_wait(bar0 + st * 8, parity);
        const uint8_t* sm = ring + (size_t)st * g.stage_bytes;
        const uint8_t* rowm[R];
        const uint8_t* rowa[R];
        bool valid[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int rl = j * groups + sr;
            valid[j] = r0 + rl < w.M;
            if (g.contig) {
                rowm[j] = sm + (size_t)rl * g.row_main;
                uintptr_t src0 = reinterpret_cast<uintptr_t>(w.aux + (size_t)r0 * g.row_aux);
                rowa[j] = sm + g.stage_main + (src0 & 15) + (size_t)rl * g.row_aux;
            } else {
                rowm[j] = sm + (size_t)rl * g.slab_main_cap;
                uintptr_t src = reinterpret_cast<uintptr_t>(w.aux + (size_t)(r0 + rl) * g.row_aux + stream_aux_bytes(w.type, s * g.slab_chunks));
                rowa[j] = sm + g.stage_main + (size_t)rl * g.slab_aux_cap + (src & 15);
            }
        }
        const int nch = min(g.slab_chunks, g.C - s * g.slab_chunks);
        for (int i = 0; i < g.cpl; i++) {
            const int cl = lr + g.lpr * i;
            const bool act = cl < nch;
            constexpr bool kShfl = TYPE == kQ4_K || TYPE == kQ5_K;  // these exchange scales by warp shuffle: no lane may skip
            if (!kShfl && !act) continue;
            const int clc = act ? cl : (cl & 7);                    // idle lanes redo an in-bounds chunk, result dropped
            const int c = s * g.slab_chunks + clc;
            const ulonglong2* xp = reinterpret_cast<const ulonglong2*>(xs + (c << 5));
            uint64_t xv[16];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                ulonglong2 t = xp[j ^ (c & 7)];
                xv[2 * j] = t.x;
                xv[2 * j + 1] = t.y;
            }
            float2 xsm = make_float2(0.0f, 0.0f);
            if (kShfl) xsm = xsum[c];
#pragma unroll
            for (int j = 0; j < R; j++) {
                float v = chunk_dot<TYPE>(rowm[j], rowa[j], clc, xv, xsm);
                if (act && valid[j]) acc[j] += v;
            }
        }
        __syncwarp();
        if (issued < nq) {  // refill the stage just drained (generic-proxy reads ordered before the async-proxy write)
            fence_proxy_async();
            issue_next();
        }
        if (s == g.n_slabs - 1) {
#pragma unroll
            for (int j = 0; j < R; j++) {
                float v = acc[j];
                for (int o = g.lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lr == 0 && valid[j]) y[r0 + j * groups + sr] = v;
            }
        }
        if (++s == g.n_slabs) { s = 0; ti++; }
    }
}

// ---- host side -----------------------------------------------------------------
int pick_lpr(int C, bool kquant) {
    int best = 32;
    double best_eff = 0.0;
    const int cand[4] = {32, 16, 8, 4};
    for (int i = 0; i < 4; i++) {
        int l = cand[i];
        int ct = (C + l - 1) / l;
        double eff = (double)C / ((double)ct * l);
        if (eff >= 0.9) return l;  // largest lane count that wastes < 10 %
        if (eff > best_eff) { best_eff = eff; best = l; }
    }
    (void)kquant;
    return best;
}

// Returns false when the shape does not fit the streamed kernel (caller falls back).
bool make_geom(int type, int M, int K, int R, bool want_contig, SGeom& g) {
    const bool kq = stream_is_kquant(type);
    if (K % (kq ? 256 : 32) || M <= 0) return false;
    g.C = K / 32;
    g.lpr = pick_lpr(g.C, kq);
    g.rows_pass = (32 / g.lpr) * R;
    g.row_main = stream_main_bytes(type, g.C);
    g.row_aux = stream_aux_bytes(type, g.C);
    const int xbytes = ((K * 4 + 127) & ~127);
    const int xsum_bytes = (type == kQ4_K || type == kQ5_K) ? ((g.C * 8 + 127) & ~127) : 0;
    int ring_budget = kSmemBudget - xbytes - xsum_bytes - 512;
    if (ring_budget > kRingBudget) ring_budget = kRingBudget;
    if (ring_budget < 16 * 1024) return false;
    const int warp_budget = ring_budget / kSWarps;
    const int ct = (g.C + g.lpr - 1) / g.lpr;  // chunks per lane per row
    // contiguous whole-row tiles when >= 2 of them fit in a warp's ring
    int tile_main = g.rows_pass * g.row_main;
    int tile_aux = g.row_aux ? ((g.rows_pass * g.row_aux + 16 + 15) & ~15) + 16 : 0;
    if (want_contig) {
        if (2 * (tile_main + tile_aux) > warp_budget) return false;
        g.contig = 1;
        g.cpl = ct;
        g.slab_chunks = g.lpr * ct;
        g.n_slabs = 1;
        g.slab_main = g.row_main;
        g.slab_main_cap = g.row_main;
        g.slab_aux_cap = 0;
        g.stage_main = tile_main;
        g.stage_bytes = (tile_main + tile_aux + 15) & ~15;
    } else {
        // K-slabs: the largest whole number of chunks per lane such that >= 3 stages fit
        const int unit = kq ? 8 : 1;  // slabs must hold whole super-blocks
        int cpl = 0;
        for (int c = ct; c >= 1; c--) {
            int sc = g.lpr * c;
            if (sc % unit) continue;
            int sm_ = stream_main_bytes(type, sc);
            int sa = stream_aux_bytes(type, sc);
            int sac = sa ? ((sa + 16 + 15) & ~15) : 0;
            if (3 * g.rows_pass * (sm_ + sac) <= warp_budget) { cpl = c; break; }
        }
        if (!cpl) return false;
        g.contig = 0;
        g.cpl = cpl;
        g.slab_chunks = g.lpr * cpl;
        g.n_slabs = (g.C + g.slab_chunks - 1) / g.slab_chunks;
        g.slab_main = stream_main_bytes(type, g.slab_chunks);
        g.slab_main_cap = g.slab_main;
        int sa = stream_aux_bytes(type, g.slab_chunks);
        g.slab_aux_cap = sa ? ((sa + 16 + 15) & ~15) : 0;
        g.stage_main = g.rows_pass * g.slab_main_cap;
        g.stage_bytes = g.stage_main + g.rows_pass * g.slab_aux_cap;
    }
    g.stages = warp_budget / g.stage_bytes;
    if (g.stages > kMaxStages) g.stages = kMaxStages;
    if (g.stages < 2) return false;
    g.n_tiles = (M + g.rows_pass - 1) / g.rows_pass;
    g.xsum_off = xbytes;
    g.ring_off = xbytes + xsum_bytes;
    g.bar_off = g.ring_off + kSWarps * g.stages * g.stage_bytes;
    g.bar_off = (g.bar_off + 15) & ~15;
    g.smem_bytes = g.bar_off + kSWarps * kMaxStages * 8;
    return g.smem_bytes <= kSmemBudget + 4096;
}

template <int TYPE, int R>
cudaError_t launch_t(const StreamW& w, const SGeom& g, const Prologue& p, float* y, const Indirect& ind, int nsel, bool pdl, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemv_stream_kernel<TYPE, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 4096);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    int ctas = (g.n_tiles + kSWarps - 1) / kSWarps;
    int cap = ZB_SMS;
    if (g.n_tiles >= 4 * 2 * ZB_SMS * kSWarps) cap = 2 * ZB_SMS;  // big streaming matrices (lm_head): both CTA slots of every SM
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
    // Tile shape: whole-row contiguous tiles with the most rows per x-chunk load that fit, K-slabs for long rows;
    // among the feasible shapes the first that still gives every warp slot of the chip a tile wins.
    static const struct { int R; bool contig; } order[6] = {{4, true}, {2, true}, {4, false}, {1, true}, {2, false}, {1, false}};
    const int slots = ZB_SMS * kSWarps;
    SGeom g{}, best{};
    int bestR = 0;
    for (int i = 0; i < 6; i++) {
        if (!make_geom(TYPE, w.M, w.K, order[i].R, order[i].contig, g)) continue;
        if (!bestR || g.rows_pass < best.rows_pass) { best = g; bestR = order[i].R; }
        if (g.n_tiles >= slots) { best = g; bestR = order[i].R; break; }
    }
    if (!bestR) return cudaErrorInvalidConfiguration;
    switch (bestR) {
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
    switch (qtype) {
        case zb::kQ4_0:
            for (int64_t b = 0; b < n32; b++) { memcpy(a + b * 2, src + b * 18, 2); memcpy(m + b * 16, src + b * 18 + 2, 16); }
            return 0;
        case zb::kQ8_0:
            for (int64_t b = 0; b < n32; b++) { memcpy(a + b * 2, src + b * 34, 2); memcpy(m + b * 32, src + b * 34 + 2, 32); }
            return 0;
        case zb::kQ4_K: memcpy(m, src, (size_t)n256 * 144); return 0;
        case zb::kQ5_K: memcpy(m, src, (size_t)n256 * 176); return 0;
        case zb::kQ6_K:
            for (int64_t b = 0; b < n256; b++) { memcpy(m + b * 208, src + b * 210, 208); memcpy(a + b * 2, src + b * 210 + 208, 2); }
            return 0;
    }
    return cudaErrorInvalidValue;
}

// 0 when a [rows, cols] matrix of this type can be run by the streamed kernel.
ZB_API int zb_stream_check(int qtype, int rows, int cols) {
    if (zb::stream_main_per8(qtype) == 0) return cudaErrorInvalidValue;
    SGeom g{};
    for (int r = 4; r >= 1; r >>= 1)
        if (make_geom(qtype, rows, cols, r, true, g) || make_geom(qtype, rows, cols, r, false, g)) return 0;
    return cudaErrorInvalidConfiguration;
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
    return launch_any(sw, pr, y, ind, nsel, (flags & 1) != 0, (cudaStream_t)stream);
}
