// Batched dequant-GEMM on the 5th-generation tensor cores (sm_100a): tcgen05.mma with the
// accumulator in tensor memory, bf16 weight tiles dequantised on the fly.
//
//   Y[T, N] = X[T, K] . deq(W[N, K])^T          T = tokens of the decode batch / prefill chunk
//
// Replaces the reference's batch > 1 path: gemm_q4_kernel N > 1 (gemm_q4.cu:116-159, one thread per
// output, no tensor cores) and dequant_q4k_f32 + cuBLAS SGEMM (dequant_q4k.cu:1-8), i.e. SURVEY 8a row a6.
//
// Mapping.  The weight rows are the UMMA M dimension (128 per CTA), the tokens the UMMA N dimension
// (16..256 per CTA), K is walked in 64-weight steps -- one "unit" of a K-quant super-block per row:
//   * 8 producer warps: thread (row, half) reads 16 bytes of quants of its row straight from the
//     stream layout (16-B aligned rows), dequantises 32 weights in f32 exactly as the bit-exact
//     dequant does (d*sc*q - dmin*m, two roundings), rounds once to bf16 and stores four 16-byte
//     chunks into the A tile in the canonical K-major SWIZZLE_128B layout; the same warps copy the
//     matching 64-column slice of the bf16 activations (hi, and lo = x - hi for small T) into the B tile;
//   * 1 MMA warp: one elected lane issues 4 (x2 with the lo part) tcgen05.mma.kind::f16 per step
//     (M=128, N=T, K=16), tcgen05.commit releases the stage; a 3-4 stage mbarrier ring couples the two;
//   * epilogue: the producer warps read the 128 x T f32 accumulator out of TMEM (tcgen05.ld) and
//     store / atomically add it to Y (split-K over CTAs when N/128 alone cannot fill 148 SMs).
// K order inside a 64-step is whatever is cheapest to dequantise; the activation pre-pass
// (zb_gemm_tc_prep_x) writes X in the same order, the contraction does not care.
//
// At decode batch sizes the kernel is bound by HBM (weights) and by the dequant instruction stream,
// not by the tensor pipe: B=32 Q5_K has 93 flop/B against a ridge of ~210.
#include <cuda_bf16.h>

#include "zb_stream.cuh"
#include "zb200.h"

namespace {

using namespace zb;

constexpr int kTM = 128;           // weight rows per CTA (UMMA M)
constexpr int kProdWarps = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kThreads = kProdThreads + 32;  // + MMA warp
constexpr int kATile = kTM * 128;  // 128 rows x 64 bf16

// ---- tcgen05 / TMEM PTX -------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (one 128-byte swizzle atom along K, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, dense
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16_2(uint64_t v) {  // two f32 -> bf16x2 (round to nearest even), low element first
    float a, b;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return pack_bf16(a, b);
}
// four masked bytes (each < 128) -> bf16x2 x 2 of (byte - off) * scale - sub, evaluated as fmul then fsub in f32 exactly like the
// bit-exact dequant (zb_quant.cuh) with packed f32x2 instructions; nbias = -(128 + off), nsub = -sub
__device__ __forceinline__ void deq4(uint32_t m, uint64_t nbias, uint64_t scale, uint64_t nsub, uint32_t& o0, uint32_t& o1) {
    uint64_t p0 = add2(pack2u(__byte_perm(m, 0x43000000u, 0x7044), __byte_perm(m, 0x43000000u, 0x7144)), nbias);
    uint64_t p1 = add2(pack2u(__byte_perm(m, 0x43000000u, 0x7244), __byte_perm(m, 0x43000000u, 0x7344)), nbias);
    o0 = pack_bf16_2(add2(mul2(p0, scale), nsub));
    o1 = pack_bf16_2(add2(mul2(p1, scale), nsub));
}

// four bytes (each < 128) of `m` as exact floats, minus bias
__device__ __forceinline__ void bytes4(uint32_t m, float bias, float (&f)[4]) {
    f[0] = __uint_as_float(__byte_perm(m, 0x43000000u, 0x7044)) - bias;
    f[1] = __uint_as_float(__byte_perm(m, 0x43000000u, 0x7144)) - bias;
    f[2] = __uint_as_float(__byte_perm(m, 0x43000000u, 0x7244)) - bias;
    f[3] = __uint_as_float(__byte_perm(m, 0x43000000u, 0x7344)) - bias;
}

// K position (inside its 256-block) of k-slot (unit, chunk i, element j) -- the order the producers emit
__host__ __device__ inline int kslot_to_k(int type, int unit, int i, int j) {
    if (type == kQ6_K) {
        int half = unit >> 1, lh = unit & 1, hi = i >> 2, w = i & 3;
        return half * 128 + hi * 64 + (j >> 2) * 32 + lh * 16 + 4 * w + (j & 3);
    }
    return unit * 64 + (j >> 2) * 32 + 4 * i + (j & 3);  // Q4_K / Q5_K: word i of the 32-byte group; low nibbles then high nibbles
}

// Element index of k-slot s of token `tok` in the activation tile image: [k-step = s/64][token][64 bf16], the eight 16-byte
// chunks of a token's 128-byte row XOR-swizzled by (token & 7) -- exactly the SWIZZLE_128B shared-memory image of a B tile whose
// first token is a multiple of 8, so a tile of consecutive tokens is one contiguous bulk copy.  ldx = token rows per k-step plane.
__host__ __device__ inline size_t ximg_index(int s, int tok, int ldx) {
    return ((size_t)(s >> 6) * ldx + tok) * 64 + (size_t)(((((s >> 3) & 7) ^ (tok & 7)) << 3) + (s & 7));
}

struct GemmArgs {
    StreamW w;
    const __nv_bfloat16* xhi;
    const __nv_bfloat16* xlo;   // nullptr: single bf16 activation
    float* y;
    int T, ldx, ldy;            // tokens, leading dims (elements)
    int nt;                     // tokens per CTA (UMMA N), multiple of 16
    int ksplit, steps_per_split;
    int stages, stage_bytes, tmem_cols;
};

template <int TYPE>
__global__ void __launch_bounds__(kThreads, 2) gemm_tc_kernel(const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * 8 + 1];
    // SWIZZLE_128B tiles must start on a 1024-byte boundary of the shared window; the launch reserves the slack
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kTM, tok0 = blockIdx.y * g.nt;
    const int step0 = blockIdx.z * g.steps_per_split;
    const int total_steps = g.w.K / 64;
    const int nsteps = min(g.steps_per_split, total_steps - step0);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), accbar = smem_u32(&bars[16]);
    const int xbytes = g.nt * 128;
    const bool split_x = g.xlo != nullptr;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; s++) {
            mbar_init(full0 + 8 * s, kProdWarps + 1);  // 8 producer warps + the thread that posts the activation bulk copy
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accbar, 1);
        fence_barrier_init();
    }
    if (warp == kProdWarps) tmem_alloc(smem_u32(&s_tmem), (uint32_t)g.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = s_tmem;

    if (warp < kProdWarps) {
        // ================= producers: dequantise A, copy B =================
        const int t = threadIdx.x, r = t & 127, hh = t >> 7;
        const int row = min(row0 + r, g.w.M - 1);  // rows past N repeat the last row; their outputs are never stored
        constexpr int BB = TYPE == kQ4_K ? 144 : (TYPE == kQ5_K ? 176 : 208);
        const uint8_t* wrow = g.w.main + (size_t)row * (g.w.K / 256) * BB;
        const uint16_t* arow = TYPE == kQ6_K ? reinterpret_cast<const uint16_t*>(g.w.aux) + (size_t)row * (g.w.K / 256) : nullptr;
        const uint32_t a_off = (uint32_t)r * 128, sw = (uint32_t)(r & 7);
        int st = 0;
        uint32_t ph = 0;
        // Global loads run one K-step ahead of the dequantisation (register double buffering): with a handful of
        // producer warps per SM the load latency is otherwise fully exposed (ncu: long_scoreboard on the first use).
        struct Pre { uint4 q, q2, hdr; float d6; };
        // activations: the prologue kernels write X as ready-made SWIZZLE_128B tile images ([k-step][token][128 B]), so the B tile of a
        // step is one contiguous span -> one bulk copy (two with the lo part) posted by a single thread, no LSU work in the producers
        const int xrows = min(g.nt, g.ldx - tok0);
        const uint32_t xcopy = (uint32_t)xrows * 128u;
        auto prefetch = [&](int step, Pre& r) {
            const int sb = step >> 2, unit = step & 3;
            const uint8_t* blk = wrow + (size_t)sb * BB;
            r.q2 = make_uint4(0, 0, 0, 0);
            r.d6 = 0.0f;
            if (TYPE == kQ6_K) {
                const int half = unit >> 1, lh = unit & 1;
                r.q = ldg128(blk + half * 64 + lh * 16);          // ql: low nibbles -> q1, high -> q3
                r.q2 = ldg128(blk + half * 64 + 32 + lh * 16);    // ql: low nibbles -> q2, high -> q4
                r.hdr = ldg128(blk + 128 + half * 32 + lh * 16);  // qh
                r.d6 = h2f(__ldg(arow + sb));
            } else {
                r.hdr = ldg128(blk);
                r.q = ldg128(blk + 16 + unit * 32 + hh * 16);
                if (TYPE == kQ5_K) r.q2 = ldg128(blk + 144 + hh * 16);
            }
            {   // pull the weights of a step 8 ahead from HBM into L2: the register prefetch above then only sees L2 latency
                const int fs = step + 8;
                if (fs < step0 + nsteps) {
                    const uint8_t* fb = wrow + (size_t)(fs >> 2) * BB + (TYPE == kQ6_K ? ((fs & 3) >> 1) * 64 + (fs & 1) * 16 : 16 + (fs & 3) * 32 + hh * 16);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(fb));
                    if (TYPE == kQ6_K) asm volatile("prefetch.global.L2 [%0];" ::"l"(fb + 128));
                }
            }
        };
        Pre cur, nxt;
        if (nsteps > 0) prefetch(step0, cur);
        for (int it = 0; it < nsteps; it++) {
            const int step = step0 + it, sb = step >> 2, unit = step & 3;
            const uint8_t* blk = wrow + (size_t)sb * BB;
            if (it + 1 < nsteps) prefetch(step + 1, nxt);
            const uint4 q = cur.q, q2 = cur.q2, hdr = cur.hdr;
            const float d6 = cur.d6;
            mbar_wait(empty0 + 8 * st, ph ^ 1u);
            uint8_t* stage = smem + (size_t)st * g.stage_bytes;
            if (t == 0) {
                const uint32_t fb = full0 + 8 * st, dst = smem_u32(stage) + kATile;
                mbar_expect_tx(fb, split_x ? 2 * xcopy : xcopy);
                bulk_g2s(dst, g.xhi + ((size_t)step * g.ldx + tok0) * 64, xcopy, fb);
                if (split_x) bulk_g2s(dst + xbytes, g.xlo + ((size_t)step * g.ldx + tok0) * 64, xcopy, fb);
            }
            (void)sb;
            // ---- dequantise 32 weights of (row, unit) -> 4 chunks of 8 bf16
            uint32_t w4[4] = {q.x, q.y, q.z, q.w};
            if (TYPE == kQ6_K) {
                const int half = unit >> 1, lh = unit & 1;
                const int8_t* scp = reinterpret_cast<const int8_t*>(blk + 192) + half * 8 + lh;
                // this thread's two scales: hh = 0 -> (q1, q2), hh = 1 -> (q3, q4)
                float s1 = d6 * (float)__ldg(scp + 4 * hh), s2 = d6 * (float)__ldg(scp + 4 * hh + 2);
                uint32_t b4[4] = {q2.x, q2.y, q2.z, q2.w}, h4[4] = {hdr.x, hdr.y, hdr.z, hdr.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t qa = ((w4[i] >> (4 * hh)) & 0x0F0F0F0Fu) | (((h4[i] >> (4 * hh)) & 0x03030303u) << 4);
                    uint32_t qb = ((b4[i] >> (4 * hh)) & 0x0F0F0F0Fu) | (((h4[i] >> (4 * hh + 2)) & 0x03030303u) << 4);
                    uint4 o;  // 128 (float trick) + 32 (Q6_K offset); d*sc*q has no subtrahend (adding -0.0f keeps the product's bits)
                    deq4(qa, pack2(-160.0f, -160.0f), pack2(s1, s1), pack2(-0.0f, -0.0f), o.x, o.y);
                    deq4(qb, pack2(-160.0f, -160.0f), pack2(s2, s2), pack2(-0.0f, -0.0f), o.z, o.w);
                    const uint32_t chunk = (uint32_t)(hh * 4 + i);
                    *reinterpret_cast<uint4*>(stage + a_off + ((chunk ^ sw) << 4)) = o;
                }
            } else {
                const int gq = unit;  // group of the super-block
                const int sh = (gq & 1) * 16;
                float d = h2f((uint16_t)(hdr.x & 0xFFFFu)), dmin = h2f((uint16_t)(hdr.x >> 16));
                uint32_t h0 = (hdr.y >> sh) & 0xFFFFu, h1 = (hdr.z >> sh) & 0xFFFFu, h2 = (hdr.w >> sh) & 0xFFFFu;
                uint32_t sc2 = gq < 2 ? (h0 & 0x3F3Fu) : ((h2 & 0x0F0Fu) | ((h0 >> 2) & 0x3030u));
                uint32_t mn2 = gq < 2 ? (h1 & 0x3F3Fu) : (((h2 >> 4) & 0x0F0Fu) | ((h1 >> 2) & 0x3030u));
                const float dsA = d * (float)(sc2 & 0xFFu), dsB = d * (float)(sc2 >> 8);
                const float dmA = dmin * (float)(mn2 & 0xFFu), dmB = dmin * (float)(mn2 >> 8);
                uint32_t hb[4] = {q2.x >> (2 * gq), q2.y >> (2 * gq), q2.z >> (2 * gq), q2.w >> (2 * gq)};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t lo = w4[i] & 0x0F0F0F0Fu, hi = (w4[i] >> 4) & 0x0F0F0F0Fu;
                    if (TYPE == kQ5_K) {
                        lo |= (hb[i] & 0x01010101u) << 4;
                        hi |= (hb[i] & 0x02020202u) << 3;
                    }
                    uint4 o;  // d*sc*q - dmin*m with the two roundings of the bit-exact dequant (zb_quant.cuh), then one rounding to bf16
                    deq4(lo, pack2(-128.0f, -128.0f), pack2(dsA, dsA), pack2(-dmA, -dmA), o.x, o.y);
                    deq4(hi, pack2(-128.0f, -128.0f), pack2(dsB, dsB), pack2(-dmB, -dmB), o.z, o.w);
                    const uint32_t chunk = (uint32_t)(hh * 4 + i);
                    *reinterpret_cast<uint4*>(stage + a_off + ((chunk ^ sw) << 4)) = o;
                }
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * st);
            if (++st == g.stages) { st = 0; ph ^= 1u; }
            cur = nxt;
        }
        // ================= epilogue: TMEM -> Y =================
        mbar_wait(accbar, 0);
        tc_fence_after();
        const int q4 = warp & 3, colhalf = warp >> 2;
        const int orow = row0 + q4 * 32 + lane;
        const int cols_per = g.nt / 2;  // nt is a multiple of 16: each half is a multiple of 8
        for (int c = colhalf * cols_per; c < (colhalf + 1) * cols_per; c += 8) {
            float v[8];
            tmem_ld8(tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c, v);
            if (orow < g.w.M) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int tok = tok0 + c + i;
                    if (tok < g.T) {
                        float* dst = g.y + (size_t)tok * g.ldy + orow;
                        if (g.ksplit > 1) atomicAdd(dst, v[i]);
                        else *dst = v[i];
                    }
                }
            }
        }
        tc_fence_before();
    } else {
        // ================= MMA issuer =================
        const uint32_t idesc = idesc_bf16(kTM, g.nt);
        int st = 0;
        uint32_t ph = 0;
        for (int it = 0; it < nsteps; it++) {
            mbar_wait(full0 + 8 * st, ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + (size_t)st * g.stage_bytes);
                const uint64_t ad = smem_desc_sw128(sa), bd = smem_desc_sw128(sa + kATile), bl = smem_desc_sw128(sa + kATile + xbytes);
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {  // 4 x (K = 16 bf16 = 32 bytes): advance the start address inside the swizzle atom
                    umma_bf16(tmem_d, ad + 2 * kk, bd + 2 * kk, idesc, (it | kk) ? 1u : 0u);
                    if (split_x) umma_bf16(tmem_d, ad + 2 * kk, bl + 2 * kk, idesc, 1u);
                }
                umma_commit(empty0 + 8 * st);             // stage reusable once these MMAs have read it
                if (it == nsteps - 1) umma_commit(accbar);  // accumulator complete
            }
            __syncwarp();
            if (++st == g.stages) { st = 0; ph ^= 1u; }
        }
        if (nsteps == 0 && lane == 0) mbar_arrive(accbar);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == kProdWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_d, (uint32_t)g.tmem_cols);
    }
}

// X f32 [T, K] -> bf16 hi (+ lo = x - hi) in the k-slot order of the format, rows padded to a multiple of 16 with zeros by the caller's allocation
__global__ void gemm_prep_x_kernel(int type, const float* __restrict__ x, int T, int K, int ldx_in, __nv_bfloat16* __restrict__ xhi,
                                   __nv_bfloat16* __restrict__ xlo, int ldx_out) {
    const int tok = blockIdx.y;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < K; s += gridDim.x * blockDim.x) {
        const int sb = s >> 8, unit = (s >> 6) & 3, i = (s >> 3) & 7, j = s & 7;
        const int k = (sb << 8) + kslot_to_k(type, unit, i, j);
        float v = x[(size_t)tok * ldx_in + k];
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        xhi[ximg_index(s, tok, ldx_out)] = h;
        if (xlo) xlo[ximg_index(s, tok, ldx_out)] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// Batched prologue: one CTA per token row.  v = a; [v = rmsnorm(v, w1)]; [v += r; sum_out = v]; [x = rmsnorm(v, w2)]
// or x = SwiGLU(gate, up); x is written as bf16 hi (+ lo) in the GEMM's k-slot order (and optionally as f32).
// Same arithmetic as the batch-1 prologue of gemv_stream.cu (rmsnorm_generic.go:10-23, silu_generic.go:22-31).
struct PrepArgs {
    const float* a; const float* r; const float* w1; const float* w2;
    float* sum_out; float* x_f32;
    __nv_bfloat16* xhi; __nv_bfloat16* xlo;
    int lda, ldr, ldsum, ldxf, ldx;
    float eps;
    int mode;   // 0: norm/add chain, 1: SwiGLU over interleaved (gate_i, up_i) pairs, 2: SwiGLU over [gate | up] halves
    int K, qtype;
};

__device__ __forceinline__ float prep_inv_rms(float sumsq, int D, float eps) { return (float)(1.0 / sqrt((double)(sumsq / (float)D + eps))); }

__global__ void __launch_bounds__(256) gemm_prep_rows_kernel(const PrepArgs p) {
    extern __shared__ float sv[];
    __shared__ float red[32];
    const int tok = blockIdx.x, tid = threadIdx.x, K = p.K;
    const float* a = p.a + (size_t)tok * p.lda;
    if (p.mode != 0) {
        for (int i = tid; i < K; i += 256) {
            float gte = p.mode == 1 ? a[2 * i] : a[i], up = p.mode == 1 ? a[2 * i + 1] : a[K + i];
            double gv = (double)gte;
            sv[i] = (float)(gv * (1.0 / (1.0 + exp(-gv)))) * up;
        }
    } else {
        const float* r = p.r ? p.r + (size_t)tok * p.ldr : nullptr;
        float* so = p.sum_out ? p.sum_out + (size_t)tok * p.ldsum : nullptr;
        float ss = 0.0f;
        for (int i = tid; i < K; i += 256) {
            float v = a[i];
            if (!p.w1 && r) {
                v = v + r[i];
                if (so) so[i] = v;
            }
            sv[i] = v;
            ss = fmaf(v, v, ss);
        }
        if (p.w1) {
            float s1 = prep_inv_rms(block_sum(ss, red), K, p.eps);
            ss = 0.0f;
            for (int i = tid; i < K; i += 256) {
                float v = sv[i] * s1 * p.w1[i];
                if (r) {
                    v = v + r[i];
                    if (so) so[i] = v;
                }
                sv[i] = v;
                ss = fmaf(v, v, ss);
            }
        }
        if (p.w2) {
            float s2 = prep_inv_rms(block_sum(ss, red), K, p.eps);
            for (int i = tid; i < K; i += 256) sv[i] = sv[i] * s2 * p.w2[i];
        }
    }
    __syncthreads();
    if (p.x_f32)
        for (int i = tid; i < K; i += 256) p.x_f32[(size_t)tok * p.ldxf + i] = sv[i];
    if (p.xhi) {
        for (int s = tid; s < K; s += 256) {
            const int sb = s >> 8, unit = (s >> 6) & 3, i = (s >> 3) & 7, j = s & 7;
            float v = sv[(sb << 8) + kslot_to_k(p.qtype, unit, i, j)];
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            p.xhi[ximg_index(s, tok, p.ldx)] = h;
            if (p.xlo) p.xlo[ximg_index(s, tok, p.ldx)] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}

// SwiGLU modes of the batched prologue need no row reduction: one thread per 4 outputs over the whole [tokens, K] slab
__global__ void __launch_bounds__(256) gemm_prep_swiglu_kernel(const PrepArgs p, int tokens) {
    const int K = p.K, per_row = K >> 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < tokens * per_row; idx += gridDim.x * blockDim.x) {
        const int tok = idx / per_row, s0 = (idx - tok * per_row) << 2;
        const float* a = p.a + (size_t)tok * p.lda;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int s = s0 + e;
            const int sb = s >> 8, unit = (s >> 6) & 3, i = (s >> 3) & 7, j = s & 7;
            const int k = p.xhi ? (sb << 8) + kslot_to_k(p.qtype, unit, i, j) : s;
            float gte = p.mode == 1 ? a[2 * k] : a[k], up = p.mode == 1 ? a[2 * k + 1] : a[K + k];
            double gv = (double)gte;
            v[e] = (float)(gv * (1.0 / (1.0 + exp(-gv)))) * up;
        }
        if (p.xhi) {
            __nv_bfloat16 h[4];
#pragma unroll
            for (int e = 0; e < 4; e++) h[e] = __float2bfloat16_rn(v[e]);
            uint2 hv;
            hv.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
            hv.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
            *reinterpret_cast<uint2*>(p.xhi + ximg_index(s0, tok, p.ldx)) = hv;
            if (p.xlo) {
                uint2 lv;
                lv.x = pack_bf16(v[0] - __bfloat162float(h[0]), v[1] - __bfloat162float(h[1]));
                lv.y = pack_bf16(v[2] - __bfloat162float(h[2]), v[3] - __bfloat162float(h[3]));
                *reinterpret_cast<uint2*>(p.xlo + ximg_index(s0, tok, p.ldx)) = lv;
            }
        }
        if (p.x_f32 && !p.xhi)
#pragma unroll
            for (int e = 0; e < 4; e++) p.x_f32[(size_t)tok * p.ldxf + s0 + e] = v[e];
    }
}

template <int TYPE>
cudaError_t launch_gemm(const GemmArgs& g, dim3 grid, size_t smem, cudaStream_t stream) {
    static DeviceOnce once;
    if (cudaError_t e = once.ensure(smem, [&] { return cudaFuncSetAttribute(gemm_tc_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }))
        return e;
    gemm_tc_kernel<TYPE><<<grid, kThreads, smem, stream>>>(g);
    return cudaGetLastError();
}

}  // namespace

// ===========================================================================
// C ABI (include/zb200.h)
// ===========================================================================
ZB_API int zb_gemm_tc_prep_x(int qtype, const float* x, int tokens, int K, int ldx, void* xhi, void* xlo, int ld_out, zb_stream_t stream) {
    if (!(qtype == zb::kQ4_K || qtype == zb::kQ5_K || qtype == zb::kQ6_K) || K % 256 || tokens <= 0 || ld_out < tokens || ld_out % 16)
        return cudaErrorInvalidValue;
    dim3 grid((K + 255) / 256, tokens);
    gemm_prep_x_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(qtype, x, tokens, K, ldx, static_cast<__nv_bfloat16*>(xhi),
                                                               static_cast<__nv_bfloat16*>(xlo), ld_out);
    return cudaGetLastError();
}

ZB_API int zb_gemm_tc_prep_rows(const zb_prep_args* a, int tokens, zb_stream_t stream) {
    if (!a || tokens <= 0 || a->K <= 0 || a->K % 256 ||
        (a->xhi && (!(a->qtype == zb::kQ4_K || a->qtype == zb::kQ5_K || a->qtype == zb::kQ6_K) || a->ldx < tokens || a->ldx % 16)))
        return cudaErrorInvalidValue;
    PrepArgs p{a->a, a->r, a->w1, a->w2, a->sum_out, a->x_f32, static_cast<__nv_bfloat16*>(a->xhi), static_cast<__nv_bfloat16*>(a->xlo),
               a->lda, a->ldr, a->ldsum, a->ldxf, a->ldx, a->eps, a->mode, a->K, a->qtype};
    if (a->mode != 0 && !(a->x_f32 && a->xhi)) {
        int blocks = (tokens * (a->K >> 2) + 255) / 256;
        if (blocks > ZB_SMS * 8) blocks = ZB_SMS * 8;
        gemm_prep_swiglu_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, tokens);
        return cudaGetLastError();
    }
    size_t smem = (size_t)a->K * 4;
    static DeviceOnce once;
    if (smem > 48 * 1024)
        if (cudaError_t e = once.ensure(smem, [&] { return cudaFuncSetAttribute(gemm_prep_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }))
            return e;
    gemm_prep_rows_kernel<<<tokens, 256, smem, (cudaStream_t)stream>>>(p);
    return cudaGetLastError();
}

// Y[tokens, N] = X . deq(W)^T.  xhi / xlo come from zb_gemm_tc_prep_x (xlo may be NULL: single-bf16 activations).
ZB_API int zb_gemm_tc_f32(const zb_stream_weight* w, const void* xhi, const void* xlo, int tokens, int ldx, float* y, int ldy,
                          zb_stream_t stream) {
    if (!w || !xhi || !y || tokens <= 0 || ldx < tokens || ldx % 16) return cudaErrorInvalidValue;
    const int type = w->qtype;
    if (!(type == zb::kQ4_K || type == zb::kQ5_K || type == zb::kQ6_K) || w->cols % 256 || w->rows <= 0) return cudaErrorInvalidValue;
    GemmArgs g{};
    g.w = zb::StreamW{static_cast<const uint8_t*>(w->main), static_cast<const uint8_t*>(w->aux), type, w->rows, w->cols};
    g.xhi = static_cast<const __nv_bfloat16*>(xhi);
    g.xlo = static_cast<const __nv_bfloat16*>(xlo);
    g.y = y;
    g.T = tokens; g.ldx = ldx; g.ldy = ldy;
    int nt = ((tokens + 15) / 16) * 16;
    if (nt > 256) nt = 256;
    g.nt = nt;
    const int row_tiles = (w->rows + kTM - 1) / kTM, tok_tiles = (tokens + nt - 1) / nt, steps = w->cols / 64;
    // split K over CTAs until the grid covers the chip (partials meet in Y with atomic adds)
    int ksplit = 1;
    while (row_tiles * tok_tiles * ksplit < 2 * ZB_SMS && ksplit * 2 <= steps / 4 && steps % (ksplit * 2) == 0) ksplit *= 2;
    g.ksplit = ksplit;
    g.steps_per_split = (steps + ksplit - 1) / ksplit;
    g.stage_bytes = kATile + nt * 128 * (xlo ? 2 : 1);
    g.stage_bytes = (g.stage_bytes + 1023) & ~1023;
    // decode batches: 3 stages keep a CTA near 75 KB so that 2-3 CTAs (16-24 producer warps) share an SM -- the kernel is bound by
    // the dequant instruction stream and its load latency, not by the tensor pipe; prefill tiles take what one CTA can get
    int stages = (200 * 1024) / g.stage_bytes;
    if (stages > 8) stages = 8;
    if (nt <= 64 && stages > 3) stages = 3;
    if (stages < 2) return cudaErrorInvalidConfiguration;
    g.stages = stages;
    int cols = 32;
    while (cols < nt) cols <<= 1;
    g.tmem_cols = cols;
    const size_t smem = (size_t)stages * g.stage_bytes + 1024;
    cudaStream_t s = (cudaStream_t)stream;
    if (ksplit > 1) {
        cudaError_t e = cudaMemset2DAsync(y, (size_t)ldy * 4, 0, (size_t)w->rows * 4, tokens, s);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(row_tiles, tok_tiles, ksplit);
    switch (type) {
        case zb::kQ4_K: return launch_gemm<zb::kQ4_K>(g, grid, smem, s);
        case zb::kQ5_K: return launch_gemm<zb::kQ5_K>(g, grid, smem, s);
        default: return launch_gemm<zb::kQ6_K>(g, grid, smem, s);
    }
}
