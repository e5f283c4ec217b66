/*
 * zoracle.c -- CPU restatement of the zerfoo transformer decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under zerfoo_b200/ may include, link or
 * dlopen this file.  It is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the *checker* and as
 * the timed CPU baseline, never as a product code path.
 *
 * Why a restatement: the arithmetic of the reference hot path lives in the
 * un-vendored Go module github.com/zerfoo/ztensor v1.19.2 (reference go.mod:25)
 * and no Go toolchain exists in this image, so the reference CPU engine
 * cannot be compiled here.  Every function below cites the reference
 * file:line it follows (paths relative to /root/reference).
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - rmsnorm / rope / swiglu / silu / softmax / sdpa / gqa / ffn: pinned
 *     against the reference's PyTorch golden vectors tests/golden/layers/NAME.json
 *     (copied as data into tests/golden/ref_layers/).
 *   - Q4_0 and Q4_K dequant + GEMV: pinned against the reference's in-test
 *     restatements and deterministic generators
 *     (internal/cuda/kernels/gemm_q4_test.go:14-85, gemv_q4k_test.go:14-92).
 *   - Q5_K / Q6_K / Q8_0: "parity unpinned" -- the reference holds no golden
 *     bytes or CPU restatement for these; this file restates the block
 *     format comments of gemv_q5k.cu:7-23, gemv_q6k.cu:7-25, gemm_q8.cu:1-7
 *     and model/gguf/loader.go:140-190 (the published ggml k-quant layout).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see Makefile).
 * -ffp-contract=off matters: Go on amd64 never fuses a*b+c, so every
 * multiply and add below rounds separately, like the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ZO_API __attribute__((visibility("default")))

/* ggml tensor type ids as stored in GGUF (model/gguf/loader.go:140-190). */
enum {
    ZO_F32 = 0, ZO_F16 = 1, ZO_Q4_0 = 2, ZO_Q8_0 = 8,
    ZO_Q4_K = 12, ZO_Q5_K = 13, ZO_Q6_K = 14
};

/* ------------------------------------------------------------------ */
/* fp16 -> f32, following internal/xblas/q4dot.go:53-80 (subnormals are
 * scaled, Inf/NaN decode to 0 "for quantization").                     */
ZO_API float zo_fp16_to_f32(uint16_t bits) {
    uint32_t sign = (bits >> 15) & 1u, exp = (bits >> 10) & 0x1Fu, frac = bits & 0x3FFu;
    if (exp == 0) {
        if (frac == 0) return 0.0f;
        float f = (float)frac / 1024.0f;
        f *= 1.0f / 16384.0f;
        return sign ? -f : f;
    }
    if (exp == 31) return 0.0f;
    uint32_t b = (sign << 31) | ((exp - 15 + 127) << 23) | (frac << 13);
    float out;
    memcpy(&out, &b, 4);
    return out;
}

/* f32 -> fp16 round-to-nearest-even (used only by test-data quantizers;
 * the reference's float16.FromFloat32 lives in an absent module, so the
 * synthetic GGUFs store already-quantized blocks and nothing downstream
 * depends on this rounding mode: SURVEY 8c "parity unpinned" note).    */
ZO_API uint16_t zo_f32_to_fp16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
    uint32_t man = x & 0x7FFFFFu;
    if (((x >> 23) & 0xFF) == 0xFF) return (uint16_t)(sign | 0x7C00u | (man ? 0x200u : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7C00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - exp);
        uint32_t half = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
    return (uint16_t)(sign | half);
}

static inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* ------------------------------------------------------------------ */
/* Block geometry (model/gguf/loader.go:140-190).                      */
ZO_API int zo_block_elems(int type) {
    switch (type) {
    case ZO_F32: case ZO_F16: return 1;
    case ZO_Q4_0: case ZO_Q8_0: return 32;
    case ZO_Q4_K: case ZO_Q5_K: case ZO_Q6_K: return 256;
    }
    return 0;
}
ZO_API int zo_block_bytes(int type) {
    switch (type) {
    case ZO_F32: return 4;
    case ZO_F16: return 2;
    case ZO_Q4_0: return 18;
    case ZO_Q8_0: return 34;
    case ZO_Q4_K: return 144;
    case ZO_Q5_K: return 176;
    case ZO_Q6_K: return 210;
    }
    return 0;
}
ZO_API int64_t zo_row_bytes(int type, int64_t k) {
    int be = zo_block_elems(type);
    if (be == 0 || k % be) return -1;
    return (k / be) * zo_block_bytes(type);
}

/* ------------------------------------------------------------------ */
/* Dequantisation, one block each.                                     */

/* Q4_0: 2 B fp16 d + 16 B; element j is the low nibble of byte j, element
 * j+16 the high nibble; w = (q-8)*d.  internal/xblas/q4dot.go:10-29,
 * internal/cuda/kernels/gemm_q4_test.go:46-76.                          */
static void deq_q4_0(const uint8_t *b, float *dst) {
    float d = zo_fp16_to_f32(rd16(b));
    for (int j = 0; j < 16; j++) {
        uint8_t q = b[2 + j];
        dst[j] = (float)((int)(q & 0x0F) - 8) * d;
        dst[j + 16] = (float)((int)(q >> 4) - 8) * d;
    }
}

/* Q8_0 (GGUF on-disk): 2 B fp16 d + 32 int8; w = q*d.
 * model/gguf/loader.go:459-502, internal/cuda/kernels/gemm_q8.cu:1-7.  */
static void deq_q8_0(const uint8_t *b, float *dst) {
    float d = zo_fp16_to_f32(rd16(b));
    for (int j = 0; j < 32; j++) dst[j] = (float)(int8_t)b[2 + j] * d;
}

/* 6-bit scale/min unpack shared by Q4_K and Q5_K.
 * internal/cuda/kernels/gemv_q4k_test.go:18-27, gemv_q4k.cu:38-56.     */
static void kq_scales(const uint8_t *sc, uint8_t *scales, uint8_t *mins) {
    for (int i = 0; i < 4; i++) {
        scales[i] = sc[i] & 63;
        mins[i] = sc[4 + i] & 63;
    }
    for (int i = 0; i < 4; i++) {
        scales[4 + i] = (uint8_t)((sc[8 + i] & 0xF) | ((sc[i] >> 6) << 4));
        mins[4 + i] = (uint8_t)((sc[8 + i] >> 4) | ((sc[4 + i] >> 6) << 4));
    }
}

/* Q4_K: internal/cuda/kernels/gemv_q4k_test.go:14-46.                  */
static void deq_q4_k(const uint8_t *b, float *dst) {
    float d = zo_fp16_to_f32(rd16(b)), dmin = zo_fp16_to_f32(rd16(b + 2));
    uint8_t scales[8], mins[8];
    kq_scales(b + 4, scales, mins);
    const uint8_t *q = b + 16;
    for (int g = 0; g < 4; g++) {
        float sc0 = d * (float)scales[2 * g], mn0 = dmin * (float)mins[2 * g];
        float sc1 = d * (float)scales[2 * g + 1], mn1 = dmin * (float)mins[2 * g + 1];
        for (int l = 0; l < 32; l++) {
            uint8_t v = q[g * 32 + l];
            dst[g * 64 + l] = sc0 * (float)(v & 0xF) - mn0;
            dst[g * 64 + l + 32] = sc1 * (float)(v >> 4) - mn1;
        }
    }
}

/* Q5_K: internal/cuda/kernels/gemv_q5k.cu:7-23,120-146.                */
static void deq_q5_k(const uint8_t *b, float *dst) {
    float d = zo_fp16_to_f32(rd16(b)), dmin = zo_fp16_to_f32(rd16(b + 2));
    uint8_t scales[8], mins[8];
    kq_scales(b + 4, scales, mins);
    const uint8_t *ql = b + 16, *qh = b + 144;
    for (int g = 0; g < 4; g++) {
        float sc0 = d * (float)scales[2 * g], mn0 = dmin * (float)mins[2 * g];
        float sc1 = d * (float)scales[2 * g + 1], mn1 = dmin * (float)mins[2 * g + 1];
        uint8_t u1 = (uint8_t)(1u << (2 * g)), u2 = (uint8_t)(2u << (2 * g));
        for (int l = 0; l < 32; l++) {
            uint8_t v = ql[g * 32 + l], h = qh[l];
            int lo = (v & 0xF) | ((h & u1) ? 16 : 0);
            int hi = (v >> 4) | ((h & u2) ? 16 : 0);
            dst[g * 64 + l] = sc0 * (float)lo - mn0;
            dst[g * 64 + l + 32] = sc1 * (float)hi - mn1;
        }
    }
}

/* Q6_K: internal/cuda/kernels/gemv_q6k.cu:7-25,88-124.                 */
static void deq_q6_k(const uint8_t *b, float *dst) {
    const uint8_t *ql = b, *qh = b + 128;
    const int8_t *sc = (const int8_t *)(b + 192);
    float d = zo_fp16_to_f32(rd16(b + 208));
    for (int half = 0; half < 2; half++) {
        const uint8_t *l4 = ql + half * 64, *h2 = qh + half * 32;
        const int8_t *s = sc + half * 8;
        float *o = dst + half * 128;
        for (int l = 0; l < 32; l++) {
            int is = l / 16;
            int q1 = (int)((l4[l] & 0xF) | ((h2[l] & 3) << 4)) - 32;
            int q2 = (int)((l4[32 + l] & 0xF) | (((h2[l] >> 2) & 3) << 4)) - 32;
            int q3 = (int)((l4[l] >> 4) | (((h2[l] >> 4) & 3) << 4)) - 32;
            int q4 = (int)((l4[32 + l] >> 4) | (((h2[l] >> 6) & 3) << 4)) - 32;
            o[l] = d * (float)s[is + 0] * (float)q1;
            o[32 + l] = d * (float)s[is + 2] * (float)q2;
            o[64 + l] = d * (float)s[is + 4] * (float)q3;
            o[96 + l] = d * (float)s[is + 6] * (float)q4;
        }
    }
}

/* Dequantise n elements (n % block == 0) of `type` into dst. */
ZO_API int zo_dequant(int type, const void *src, float *dst, int64_t n) {
    int be = zo_block_elems(type), bb = zo_block_bytes(type);
    if (!be || n % be) return -1;
    const uint8_t *p = (const uint8_t *)src;
    int64_t nb = n / be;
    switch (type) {
    case ZO_F32: memcpy(dst, src, (size_t)n * 4); return 0;
    case ZO_F16: for (int64_t i = 0; i < n; i++) dst[i] = zo_fp16_to_f32(rd16(p + 2 * i)); return 0;
    default: break;
    }
#pragma omp parallel for schedule(static) if (nb > 4096)
    for (int64_t i = 0; i < nb; i++) {
        const uint8_t *b = p + i * bb;
        float *o = dst + i * be;
        switch (type) {
        case ZO_Q4_0: deq_q4_0(b, o); break;
        case ZO_Q8_0: deq_q8_0(b, o); break;
        case ZO_Q4_K: deq_q4_k(b, o); break;
        case ZO_Q5_K: deq_q5_k(b, o); break;
        case ZO_Q6_K: deq_q6_k(b, o); break;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Row dot products.                                                   */

/* Q4_0 row . x in the reference CPU order (internal/xblas/q4dot.go:10-50):
 * per block  sum = sum_p( lo_p*x[p] ; hi_p*x[p+16] ) interleaved, then
 * *scale, blocks accumulated left to right.                            */
static float dot_q4_0_row(const uint8_t *row, const float *x, int64_t nblk) {
    float total = 0.0f;
    for (int64_t bi = 0; bi < nblk; bi++) {
        const uint8_t *b = row + bi * 18;
        float scale = zo_fp16_to_f32(rd16(b));
        const float *xs = x + bi * 32;
        float sum = 0.0f;
        for (int p = 0; p < 16; p++) {
            uint8_t v = b[2 + p];
            float lo = (float)((int)(v & 0x0F) - 8);
            float hi = (float)((int)(v >> 4) - 8);
            sum += lo * xs[p];
            sum += hi * xs[p + 16];
        }
        total += sum * scale;
    }
    return total;
}

/* Generic row: dequantise (bit-exact) then sequential f32 sum in k order,
 * as the reference's in-test CPU reference does
 * (internal/cuda/kernels/gemv_q4k_test.go:82-87).                       */
static float dot_generic_row(int type, const uint8_t *row, const float *x, int64_t k, float *tmp) {
    zo_dequant(type, row, tmp, k);
    float s = 0.0f;
    for (int64_t i = 0; i < k; i++) s += tmp[i] * x[i];
    return s;
}

static double dot_generic_row_f64(int type, const uint8_t *row, const float *x, int64_t k, float *tmp) {
    zo_dequant(type, row, tmp, k);
    double s = 0.0;
    for (int64_t i = 0; i < k; i++) s += (double)tmp[i] * (double)x[i];
    return s;
}

/* y[rows] = W[rows,K] . x[K]; W in `type` blocks, row-major.  Row-parallel
 * like gemmF32Q4NTParallel (internal/xblas/gemm_quant.go:122-160): the
 * per-output order is independent of the worker count.                 */
ZO_API int zo_gemv(int type, const void *w, int64_t rows, int64_t k, const float *x, float *y) {
    int64_t rb = zo_row_bytes(type, k);
    if (rb < 0) return -1;
    const uint8_t *p = (const uint8_t *)w;
#pragma omp parallel if (rows * k >= 65536)
    {
        float *tmp = (type == ZO_Q4_0) ? NULL : (float *)malloc((size_t)k * 4);
#pragma omp for schedule(static)
        for (int64_t r = 0; r < rows; r++) {
            if (type == ZO_Q4_0) y[r] = dot_q4_0_row(p + r * rb, x, k / 32);
            else y[r] = dot_generic_row(type, p + r * rb, x, k, tmp);
        }
        free(tmp);
    }
    return 0;
}

/* Same contraction accumulated in f64: the "exact" reference used to
 * report how far both f32 orders are from the true value.              */
ZO_API int zo_gemv_f64(int type, const void *w, int64_t rows, int64_t k, const float *x, double *y) {
    int64_t rb = zo_row_bytes(type, k);
    if (rb < 0) return -1;
    const uint8_t *p = (const uint8_t *)w;
#pragma omp parallel if (rows * k >= 65536)
    {
        float *tmp = (float *)malloc((size_t)k * 4);
#pragma omp for schedule(static)
        for (int64_t r = 0; r < rows; r++) y[r] = dot_generic_row_f64(type, p + r * rb, x, k, tmp);
        free(tmp);
    }
    return 0;
}

/* C[m,n] = X[m,k] . W[n,k]^T, each element in the GEMV order above
 * (GemmF32Q4NT loops q4DotRow per (i,j): gemm_quant.go:112-119).        */
ZO_API int zo_gemm_nt(int type, const void *w, int64_t n, int64_t k, const float *x, int64_t m, float *c) {
    float *col = (float *)malloc((size_t)n * 4);
    for (int64_t i = 0; i < m; i++) {
        if (zo_gemv(type, w, n, k, x + i * k, col)) { free(col); return -1; }
        memcpy(c + i * n, col, (size_t)n * 4);
    }
    free(col);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Elementwise / normalisation ops.                                    */

/* internal/xblas/rmsnorm_generic.go:10-23. Returns the per-row scale.   */
ZO_API float zo_rmsnorm(float *out, const float *x, const float *w, int dim, float eps) {
    float ss = 0.0f;
    for (int i = 0; i < dim; i++) ss += x[i] * x[i];
    float s = (float)(1.0 / sqrt((double)(ss / (float)dim + eps)));
    for (int i = 0; i < dim; i++) out[i] = x[i] * s * w[i];
    return s;
}

/* sum = a + r; normed = rmsnorm(sum).  inference/fused_add_rmsnorm_node.go:32-47,
 * internal/cuda/kernels/fused_add_rmsnorm.cu:17-55.                     */
ZO_API void zo_add_rmsnorm(float *normed, float *sum, const float *a, const float *r,
                           const float *w, int dim, float eps) {
    for (int i = 0; i < dim; i++) sum[i] = a[i] + r[i];
    zo_rmsnorm(normed, sum, w, dim, eps);
}

/* out = rmsnorm(x,w) + r.  inference/fused_norm_add_node.go:31-49.      */
ZO_API void zo_norm_add(float *out, const float *x, const float *w, const float *r, int dim, float eps) {
    float ss = 0.0f;
    for (int i = 0; i < dim; i++) ss += x[i] * x[i];
    float s = (float)(1.0 / sqrt((double)(ss / (float)dim + eps)));
    for (int i = 0; i < dim; i++) out[i] = x[i] * s * w[i] + r[i];
}

/* internal/xblas/silu_generic.go:12-31 (f64 exp).                       */
ZO_API void zo_silu(float *out, const float *x, int n) {
    for (int i = 0; i < n; i++) {
        double v = (double)x[i];
        out[i] = (float)(v / (1.0 + exp(-v)));
    }
}
ZO_API void zo_swiglu(float *out, const float *gate, const float *up, int n) {
    for (int i = 0; i < n; i++) {
        double g = (double)gate[i];
        double sig = 1.0 / (1.0 + exp(-g));
        out[i] = (float)(g * sig) * up[i];
    }
}

/* internal/xblas/softmax_generic.go:10-27, in place.                    */
ZO_API void zo_softmax(float *s, int n) {
    float mx = s[0];
    for (int i = 1; i < n; i++) if (s[i] > mx) mx = s[i];
    float sum = 0.0f;
    for (int i = 0; i < n; i++) {
        s[i] = (float)exp((double)(s[i] - mx));
        sum += s[i];
    }
    float inv = 1.0f / sum;
    for (int i = 0; i < n; i++) s[i] *= inv;
}

/* Half-split (NeoX) RoPE with pass-through tail.
 * internal/xblas/rope_generic.go:7-19.                                  */
ZO_API void zo_rope(float *out, const float *in, const float *cs, const float *sn, int half, int head_dim) {
    for (int i = 0; i < half; i++) {
        float a = in[i], b = in[i + half];
        out[i] = a * cs[i] - b * sn[i];
        out[i + half] = b * cs[i] + a * sn[i];
    }
    for (int i = 2 * half; i < head_dim; i++) out[i] = in[i];
}

/* cos/sin tables [positions, rot/2]: inv = 1/base^(2i/rot) and the angles
 * in f64, stored as f32.
 * layers/embeddings/rotary_positional_embedding.go:117-160.             */
ZO_API void zo_rope_tables(float *cs, float *sn, int positions, int rotary_dim, double base) {
    int half = rotary_dim / 2;
    double *inv = (double *)malloc(sizeof(double) * (size_t)half);
    for (int i = 0; i < half; i++) inv[i] = 1.0 / pow(base, (double)(2 * i) / (double)rotary_dim);
    for (int p = 0; p < positions; p++)
        for (int j = 0; j < half; j++) {
            double ang = (double)p * inv[j];
            cs[(size_t)p * half + j] = (float)cos(ang);
            sn[(size_t)p * half + j] = (float)sin(ang);
        }
    free(inv);
}

/* Gemma softcap on the CPU engine: rational tanh, clamped at |x|>=4.5.
 * inference/arch_llama.go:15-27,184-213.                                */
static float tanhf32_ref(float x) {
    if (x > 4.5f) return 1.0f;
    if (x < -4.5f) return -1.0f;
    float x2 = x * x;
    return x * (27.0f + x2) / (27.0f + 9.0f * x2);
}
ZO_API void zo_softcap(float *logits, int n, float cap) {
    float inv = (float)(1.0 / (double)cap);
    for (int i = 0; i < n; i++) logits[i] = cap * tanhf32_ref(logits[i] * inv);
}

/* Greedy argmax, strict '>' so the lowest index wins ties.
 * generate/generator.go:563-572, internal/cuda/kernels/argmax.cu:40-48. */
ZO_API int zo_argmax(const float *x, int n) {
    int best = 0;
    float bv = x[0];
    for (int i = 1; i < n; i++) if (x[i] > bv) { bv = x[i]; best = i; }
    return best;
}

/* Single-query attention for one head over a contiguous cache:
 * scores = (q.K_t) * scale -> softmax -> sum_t p_t V_t.  CPU path of
 * layers/attention/scaled_dot_product_attention.go:198-340
 * (MatMulTransposeB -> MulScalar -> Softmax -> MatMul).
 * k,v: [kv_len, stride] with this head's hd values at offset 0.        */
ZO_API void zo_attn_decode_head(float *out, const float *q, const float *k, const float *v,
                                int kv_len, int hd, int64_t stride, float scale, float *scratch) {
    for (int t = 0; t < kv_len; t++) {
        const float *kt = k + (int64_t)t * stride;
        float s = 0.0f;
        for (int d = 0; d < hd; d++) s += q[d] * kt[d];
        scratch[t] = s * scale;
    }
    zo_softmax(scratch, kv_len);
    for (int d = 0; d < hd; d++) out[d] = 0.0f;
    for (int t = 0; t < kv_len; t++) {
        const float *vt = v + (int64_t)t * stride;
        float p = scratch[t];
        for (int d = 0; d < hd; d++) out[d] += p * vt[d];
    }
}

/* Causal self-attention over a whole sequence, one head, [seq, hd] rows.
 * Additive -1e9 mask (scaled_dot_product_attention.go:16-27,305-313): in
 * f32 score-1e9 == -1e9 and exp underflows to exactly 0, so masked
 * positions contribute nothing but are still part of the row.          */
ZO_API void zo_attn_causal_head(float *out, const float *q, const float *k, const float *v,
                                int seq, int hd, float scale, float *scratch) {
    for (int i = 0; i < seq; i++) {
        for (int t = 0; t < seq; t++) {
            float s = 0.0f;
            for (int d = 0; d < hd; d++) s += q[i * hd + d] * k[t * hd + d];
            s *= scale;
            if (t > i) s += -1e9f;
            scratch[t] = s;
        }
        zo_softmax(scratch, seq);
        for (int d = 0; d < hd; d++) out[i * hd + d] = 0.0f;
        for (int t = 0; t < seq; t++)
            for (int d = 0; d < hd; d++) out[i * hd + d] += scratch[t] * v[t * hd + d];
    }
}

/* ------------------------------------------------------------------ */
/* Minimal GGUF v2/v3 reader (model/gguf/parser.go:96, loader.go:86-104). */

typedef struct {
    char name[128];
    int type;
    int n_dims;
    int64_t ne[4];          /* GGML order: ne[0] innermost */
    const uint8_t *data;
} zo_tensor;

typedef struct {
    char key[128];
    int vtype;              /* 4 u32, 5 i32, 6 f32, 8 string, 10 u64 ... */
    double num;
    char str[128];
} zo_kv;

typedef struct {
    uint8_t *buf;
    size_t size;
    int n_kv, n_tensors;
    zo_kv *kv;
    zo_tensor *tensors;
} zo_gguf;

typedef struct { const uint8_t *p, *end; int bad; } rdr;
static uint64_t r_u64(rdr *r) { if (r->p + 8 > r->end) { r->bad = 1; return 0; } uint64_t v; memcpy(&v, r->p, 8); r->p += 8; return v; }
static uint32_t r_u32(rdr *r) { if (r->p + 4 > r->end) { r->bad = 1; return 0; } uint32_t v; memcpy(&v, r->p, 4); r->p += 4; return v; }
static void r_str(rdr *r, char *dst, size_t cap) {
    uint64_t n = r_u64(r);
    if (r->bad || r->p + n > r->end) { r->bad = 1; if (cap) dst[0] = 0; return; }
    size_t c = n < cap - 1 ? (size_t)n : cap - 1;
    memcpy(dst, r->p, c);
    dst[c] = 0;
    r->p += n;
}
static size_t scalar_size(uint32_t t) {
    switch (t) { case 0: case 1: case 7: return 1; case 2: case 3: return 2; case 4: case 5: case 6: return 4; case 10: case 11: case 12: return 8; }
    return 0;
}
static double r_scalar(rdr *r, uint32_t t) {
    size_t n = scalar_size(t);
    if (!n || r->p + n > r->end) { r->bad = 1; return 0; }
    double out = 0;
    switch (t) {
    case 0: out = *(const uint8_t *)r->p; break;
    case 1: out = *(const int8_t *)r->p; break;
    case 2: { uint16_t v; memcpy(&v, r->p, 2); out = v; } break;
    case 3: { int16_t v; memcpy(&v, r->p, 2); out = v; } break;
    case 4: { uint32_t v; memcpy(&v, r->p, 4); out = v; } break;
    case 5: { int32_t v; memcpy(&v, r->p, 4); out = v; } break;
    case 6: { float v; memcpy(&v, r->p, 4); out = v; } break;
    case 7: out = *(const uint8_t *)r->p; break;
    case 10: { uint64_t v; memcpy(&v, r->p, 8); out = (double)v; } break;
    case 11: { int64_t v; memcpy(&v, r->p, 8); out = (double)v; } break;
    case 12: { double v; memcpy(&v, r->p, 8); out = v; } break;
    }
    r->p += n;
    return out;
}

ZO_API void zo_gguf_close(zo_gguf *g) {
    if (!g) return;
    free(g->buf); free(g->kv); free(g->tensors); free(g);
}

ZO_API zo_gguf *zo_gguf_open(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    zo_gguf *g = (zo_gguf *)calloc(1, sizeof(*g));
    g->buf = (uint8_t *)malloc((size_t)sz);
    g->size = (size_t)sz;
    if (!g->buf || fread(g->buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); zo_gguf_close(g); return NULL; }
    fclose(f);
    rdr r = { g->buf, g->buf + sz, 0 };
    if (r_u32(&r) != 0x46554747u) { zo_gguf_close(g); return NULL; }
    uint32_t ver = r_u32(&r);
    if (ver < 2 || ver > 3) { zo_gguf_close(g); return NULL; }
    uint64_t nt = r_u64(&r), nkv = r_u64(&r);
    if (r.bad || nt > (1u << 20) || nkv > (1u << 20)) { zo_gguf_close(g); return NULL; }
    g->kv = (zo_kv *)calloc(nkv ? nkv : 1, sizeof(zo_kv));
    g->tensors = (zo_tensor *)calloc(nt ? nt : 1, sizeof(zo_tensor));
    g->n_kv = (int)nkv; g->n_tensors = (int)nt;
    uint64_t alignment = 32;
    for (uint64_t i = 0; i < nkv && !r.bad; i++) {
        zo_kv *kv = &g->kv[i];
        r_str(&r, kv->key, sizeof kv->key);
        uint32_t t = r_u32(&r);
        kv->vtype = (int)t;
        if (t == 8) r_str(&r, kv->str, sizeof kv->str);
        else if (t == 9) {
            uint32_t et = r_u32(&r);
            uint64_t cnt = r_u64(&r);
            for (uint64_t j = 0; j < cnt && !r.bad; j++) {
                if (et == 8) { char tmp[8]; r_str(&r, tmp, sizeof tmp); }
                else r_scalar(&r, et);
            }
            kv->num = (double)cnt;
        } else kv->num = r_scalar(&r, t);
        if (!strcmp(kv->key, "general.alignment") && kv->num > 0) alignment = (uint64_t)kv->num;
    }
    for (uint64_t i = 0; i < nt && !r.bad; i++) {
        zo_tensor *t = &g->tensors[i];
        r_str(&r, t->name, sizeof t->name);
        t->n_dims = (int)r_u32(&r);
        if (t->n_dims > 4) { r.bad = 1; break; }
        for (int d = 0; d < 4; d++) t->ne[d] = 1;
        for (int d = 0; d < t->n_dims; d++) t->ne[d] = (int64_t)r_u64(&r);
        t->type = (int)r_u32(&r);
        t->data = (const uint8_t *)(uintptr_t)r_u64(&r); /* offset for now */
    }
    if (r.bad) { zo_gguf_close(g); return NULL; }
    size_t data_start = (size_t)(r.p - g->buf);
    data_start = (data_start + alignment - 1) / alignment * alignment;
    for (int i = 0; i < g->n_tensors; i++) {
        zo_tensor *t = &g->tensors[i];
        size_t off = (size_t)(uintptr_t)t->data;
        int64_t n = t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3];
        int be = zo_block_elems(t->type);
        if (!be || n % be || data_start + off + (size_t)(n / be) * zo_block_bytes(t->type) > g->size) {
            zo_gguf_close(g);
            return NULL;
        }
        t->data = g->buf + data_start + off;
    }
    return g;
}

static const zo_kv *find_kv(const zo_gguf *g, const char *key) {
    for (int i = 0; i < g->n_kv; i++) if (!strcmp(g->kv[i].key, key)) return &g->kv[i];
    return NULL;
}
static const zo_tensor *find_tensor(const zo_gguf *g, const char *name) {
    for (int i = 0; i < g->n_tensors; i++) if (!strcmp(g->tensors[i].name, name)) return &g->tensors[i];
    return NULL;
}
ZO_API int zo_gguf_n_tensors(const zo_gguf *g) { return g->n_tensors; }
ZO_API const char *zo_gguf_tensor_name(const zo_gguf *g, int i) { return g->tensors[i].name; }
ZO_API int zo_gguf_tensor_info(const zo_gguf *g, const char *name, int *type, int64_t *rows, int64_t *cols, const void **data) {
    const zo_tensor *t = find_tensor(g, name);
    if (!t) return -1;
    *type = t->type; *cols = t->ne[0]; *rows = t->ne[1] * t->ne[2] * t->ne[3]; *data = t->data;
    return 0;
}

/* ------------------------------------------------------------------ */
/* Model: decoder-only transformer as built by buildTransformerGraph
 * (inference/arch_common.go:104-562) and buildMixtralGraph
 * (inference/arch_mixtral.go:47-335) on the CPU-default mmap path
 * (native blocks everywhere, no requant: SURVEY 0.5).                  */

typedef struct { int type; int64_t rows, cols; const uint8_t *data; } zo_w;

typedef struct {
    zo_w attn_norm, q, k, v, o, q_norm, k_norm, post_attn_norm, ffn_norm, post_ffw_norm;
    zo_w gate, up, down;
    zo_w router, gate_exps, up_exps, down_exps;   /* MoE */
    float *cos_tbl, *sin_tbl;                     /* [max_seq, hd/2] for this layer's base */
    float *kc, *vc;                               /* [max_seq, nKV*hd] */
} zo_layer;

typedef struct zo_model {
    zo_gguf *g;
    char arch[128];
    int vocab, hidden, layers, n_q, n_kv, hd, ffn, max_seq;
    int n_experts, top_k;
    float eps, softcap, embed_scale;
    int post_norm, qk_norm;
    double rope_base, rope_local_base;
    int sw_pattern;
    int kv_f16;           /* KV cache stored as fp16 (generate/tensor_cache.go:224-238): values rounded on the append */
    int prefill_window;   /* Mistral-family sliding window: applied by the PROMPT pass only (see zo_model_prefill) */
    int in_prefill;
    zo_w embed, out_norm, lm_head;
    zo_layer *L;
    float *tbl_global_cos, *tbl_global_sin, *tbl_local_cos, *tbl_local_sin;
    int pos;
    /* scratch */
    float *hid, *normed, *qkv, *attn, *proj, *res, *gate_b, *up_b, *act, *scores, *logits, *rowbuf, *moe_out;
} zo_model;

static int get_w(const zo_gguf *g, const char *name, zo_w *w, int required) {
    const zo_tensor *t = find_tensor(g, name);
    if (!t) { memset(w, 0, sizeof *w); w->type = -1; return required ? -1 : 0; }
    w->type = t->type; w->cols = t->ne[0]; w->rows = t->ne[1] * t->ne[2] * t->ne[3]; w->data = t->data;
    return 0;
}
static double kv_num(const zo_gguf *g, const char *arch, const char *suffix, double dflt) {
    char key[192];
    snprintf(key, sizeof key, "%s.%s", arch, suffix);
    const zo_kv *kv = find_kv(g, key);
    return (kv && kv->vtype != 8 && kv->vtype != 9) ? kv->num : dflt;
}

ZO_API void zo_model_free(zo_model *m) {
    if (!m) return;
    if (m->L) for (int i = 0; i < m->layers; i++) { free(m->L[i].kc); free(m->L[i].vc); }
    free(m->L);
    free(m->tbl_global_cos); free(m->tbl_global_sin); free(m->tbl_local_cos); free(m->tbl_local_sin);
    free(m->hid); free(m->normed); free(m->qkv); free(m->attn); free(m->proj); free(m->res);
    free(m->gate_b); free(m->up_b); free(m->act); free(m->scores); free(m->logits); free(m->rowbuf); free(m->moe_out);
    zo_gguf_close(m->g);
    free(m);
}

/* max_seq_override > 0 caps the KV cache / RoPE table length. */
ZO_API zo_model *zo_model_load(const char *path, int max_seq_override) {
    zo_gguf *g = zo_gguf_open(path);
    if (!g) return NULL;
    zo_model *m = (zo_model *)calloc(1, sizeof *m);
    m->g = g;
    const zo_kv *a = find_kv(g, "general.architecture");
    if (!a || a->vtype != 8) { zo_model_free(m); return NULL; }
    snprintf(m->arch, sizeof m->arch, "%s", a->str);
    const char *ar = m->arch;
    /* model/gguf/arch.go:146-245 */
    m->vocab = (int)kv_num(g, ar, "vocab_size", 0);
    m->hidden = (int)kv_num(g, ar, "embedding_length", 0);
    m->layers = (int)kv_num(g, ar, "block_count", 0);
    m->n_q = (int)kv_num(g, ar, "attention.head_count", 0);
    m->n_kv = (int)kv_num(g, ar, "attention.head_count_kv", m->n_q);
    m->ffn = (int)kv_num(g, ar, "feed_forward_length", 0);
    m->max_seq = (int)kv_num(g, ar, "context_length", 2048);
    m->rope_base = kv_num(g, ar, "rope.freq_base", 0);
    if (m->rope_base == 0) m->rope_base = kv_num(g, ar, "rope.global.freq_base", 10000.0);
    m->hd = (int)kv_num(g, ar, "attention.key_length", 0);
    if (m->hd <= 0 && m->n_q > 0) m->hd = m->hidden / m->n_q;
    m->softcap = (float)kv_num(g, ar, "final_logit_softcapping", 0);
    m->rope_local_base = kv_num(g, ar, "rope.local.freq_base", 0);
    m->sw_pattern = m->rope_local_base > 0 ? 6 : 0;
    /* arch_mistral.go:40 / arch_mixtral.go:191 / arch_starcoder2.go:42 hand cfg.SlidingWindow to every attention layer
     * (arch_common.go:333-335); the Gemma builders do not. */
    if (!strcmp(ar, "mistral") || !strcmp(ar, "mixtral") || !strcmp(ar, "starcoder2"))
        m->prefill_window = (int)kv_num(g, ar, "attention.sliding_window", 0);
    m->eps = (float)kv_num(g, ar, "attention.layer_norm_rms_epsilon", 0);
    if (!(m->eps > 0)) m->eps = 1e-5f;                     /* arch_common.go:115-118 */
    m->n_experts = (int)kv_num(g, ar, "expert_count", 0);
    m->top_k = (int)kv_num(g, ar, "expert_used_count", 0);
    if (max_seq_override > 0 && max_seq_override < m->max_seq) m->max_seq = max_seq_override;
    int is_gemma = !strncmp(ar, "gemma", 5);
    int is_gemma3 = !strcmp(ar, "gemma3");
    int is_moe = !strcmp(ar, "mixtral") || m->n_experts > 0;
    if (is_moe) {                                           /* arch_mixtral.go:78-85 */
        if (m->n_experts == 0) m->n_experts = 8;
        if (m->top_k == 0) m->top_k = 2;
    }
    if (is_gemma) m->embed_scale = (float)sqrt((double)m->hidden);   /* arch_gemma.go:38 */
    if (is_gemma3) { m->post_norm = 1; m->qk_norm = 1; } else m->softcap = 0; /* arch_gemma.go:43-47 */
    if (m->vocab <= 0 || m->hidden <= 0 || m->layers <= 0 || m->n_q <= 0 || m->hd <= 0) { zo_model_free(m); return NULL; }

    int bad = 0;
    bad |= get_w(g, "token_embd.weight", &m->embed, 1);
    bad |= get_w(g, "output_norm.weight", &m->out_norm, 1);
    get_w(g, "output.weight", &m->lm_head, 0);
    if (m->lm_head.type < 0) m->lm_head = m->embed;          /* tied head */
    if (m->vocab != m->embed.rows) m->vocab = (int)m->embed.rows;
    m->L = (zo_layer *)calloc((size_t)m->layers, sizeof(zo_layer));
    int half = m->hd / 2;
    size_t tsz = (size_t)m->max_seq * half;
    m->tbl_global_cos = (float *)malloc(tsz * 4); m->tbl_global_sin = (float *)malloc(tsz * 4);
    zo_rope_tables(m->tbl_global_cos, m->tbl_global_sin, m->max_seq, m->hd, m->rope_base);
    if (m->sw_pattern > 0) {
        m->tbl_local_cos = (float *)malloc(tsz * 4); m->tbl_local_sin = (float *)malloc(tsz * 4);
        zo_rope_tables(m->tbl_local_cos, m->tbl_local_sin, m->max_seq, m->hd, m->rope_local_base);
    }
    for (int i = 0; i < m->layers; i++) {
        zo_layer *L = &m->L[i];
        char n[160];
#define W(field, suffix, req) do { snprintf(n, sizeof n, "blk.%d." suffix ".weight", i); bad |= get_w(g, n, &L->field, req); } while (0)
        W(attn_norm, "attn_norm", 1); W(q, "attn_q", 1); W(k, "attn_k", 1); W(v, "attn_v", 1); W(o, "attn_output", 1);
        W(q_norm, "attn_q_norm", m->qk_norm); W(k_norm, "attn_k_norm", m->qk_norm);
        W(post_attn_norm, "post_attention_norm", m->post_norm);
        W(ffn_norm, "ffn_norm", 1);
        W(post_ffw_norm, "post_ffw_norm", m->post_norm);
        if (is_moe) {
            W(router, "ffn_gate_inp", 1); W(gate_exps, "ffn_gate_exps", 1); W(up_exps, "ffn_up_exps", 1); W(down_exps, "ffn_down_exps", 1);
        } else {
            L->router.type = -1;
            W(gate, "ffn_gate", 1); W(up, "ffn_up", 1); W(down, "ffn_down", 1);
        }
#undef W
        /* arch_common.go:171-178: every sw_pattern-th layer is global. */
        int global = !(m->sw_pattern > 0 && ((i + 1) % m->sw_pattern != 0));
        L->cos_tbl = global ? m->tbl_global_cos : m->tbl_local_cos;
        L->sin_tbl = global ? m->tbl_global_sin : m->tbl_local_sin;
        size_t kvsz = (size_t)m->max_seq * m->n_kv * m->hd;
        L->kc = (float *)calloc(kvsz, 4); L->vc = (float *)calloc(kvsz, 4);
        if (!L->kc || !L->vc) bad = 1;
    }
    if (bad) { zo_model_free(m); return NULL; }
    int qdim = m->n_q * m->hd, kvdim = m->n_kv * m->hd;
    int maxd = m->hidden > m->ffn ? m->hidden : m->ffn;
    if (qdim > maxd) maxd = qdim;
    m->hid = (float *)malloc((size_t)m->hidden * 4);
    m->normed = (float *)malloc((size_t)m->hidden * 4);
    m->qkv = (float *)malloc((size_t)(qdim + 2 * kvdim) * 4);
    m->attn = (float *)malloc((size_t)qdim * 4);
    m->proj = (float *)malloc((size_t)m->hidden * 4);
    m->res = (float *)malloc((size_t)m->hidden * 4);
    m->gate_b = (float *)malloc((size_t)m->ffn * 4);
    m->up_b = (float *)malloc((size_t)m->ffn * 4);
    m->act = (float *)malloc((size_t)m->ffn * 4);
    m->scores = (float *)malloc((size_t)m->max_seq * 4);
    m->logits = (float *)malloc((size_t)m->vocab * 4);
    m->rowbuf = (float *)malloc((size_t)maxd * 4);
    m->moe_out = (float *)malloc((size_t)m->hidden * 4);
    return m;
}

ZO_API void zo_model_reset(zo_model *m) { m->pos = 0; }
ZO_API int zo_model_pos(const zo_model *m) { return m->pos; }
ZO_API void zo_model_dims(const zo_model *m, int *out) {
    out[0] = m->vocab; out[1] = m->hidden; out[2] = m->layers; out[3] = m->n_q; out[4] = m->n_kv;
    out[5] = m->hd; out[6] = m->ffn; out[7] = m->max_seq; out[8] = m->n_experts; out[9] = m->top_k;
}
ZO_API const float *zo_model_logits(const zo_model *m) { return m->logits; }
ZO_API const float *zo_model_hidden(const zo_model *m) { return m->hid; }
ZO_API const float *zo_model_kcache(const zo_model *m, int layer) { return m->L[layer].kc; }
ZO_API const float *zo_model_vcache(const zo_model *m, int layer) { return m->L[layer].vc; }

static int wgemv(const zo_w *w, const float *x, float *y) { return zo_gemv(w->type, w->data, w->rows, w->cols, x, y); }

/* SwiGLU FFN with weights (gate, up, down): layers/core/ffn.go:175-265. */
static void ffn_forward(zo_model *m, const zo_w *gate, const zo_w *up, const zo_w *down, const float *x, float *out) {
    wgemv(gate, x, m->gate_b);
    wgemv(up, x, m->up_b);
    zo_swiglu(m->act, m->gate_b, m->up_b, (int)gate->rows);
    wgemv(down, m->act, out);
}

/* Expert e of a stacked [E, rows, cols] tensor, sliced at block boundaries
 * (inference/arch_mixtral.go buildExpertFFN / extractExpertSlice).       */
static zo_w expert_slice(const zo_w *w, int e, int n_experts) {
    zo_w s = *w;
    s.rows = w->rows / n_experts;
    s.data = w->data + (int64_t)e * s.rows * zo_row_bytes(w->type, w->cols);
    return s;
}

/* Router: softmax over E logits, top-k by probability (descending), weights
 * renormalised to sum 1.  layers/core/moe.go:74-146.  sort.Slice is
 * unstable on exact ties; this restatement breaks ties towards the lowest
 * index (synthetic tests avoid exact ties).  logits is overwritten with
 * the probabilities.                                                   */
ZO_API void zo_moe_route(float *logits, int E, int K, int *idx_out, float *w_out) {
    int idx[256];
    if (K > E) K = E;
    zo_softmax(logits, E);
    for (int i = 0; i < E; i++) idx[i] = i;
    for (int i = 1; i < E; i++) {
        int v = idx[i], j = i - 1;
        while (j >= 0 && logits[idx[j]] < logits[v]) { idx[j + 1] = idx[j]; j--; }
        idx[j + 1] = v;
    }
    float wsum = 0.0f;
    for (int k = 0; k < K; k++) { w_out[k] = logits[idx[k]]; wsum = wsum + w_out[k]; idx_out[k] = idx[k]; }
    for (int k = 0; k < K; k++) w_out[k] = w_out[k] / wsum;
}

/* Decode-time combine: out = sum_k w_k * FFN_{e_k}(x), experts visited in
 * descending-probability order.  layers/core/moe.go:463-485.            */
static void moe_forward(zo_model *m, const zo_layer *L, const float *x, float *out) {
    int E = m->n_experts, K = m->top_k > E ? E : m->top_k;
    float probs[256], wk[256];
    int idx[256];
    wgemv(&L->router, x, probs);
    zo_moe_route(probs, E, K, idx, wk);
    for (int i = 0; i < m->hidden; i++) out[i] = 0.0f;
    for (int k = 0; k < K; k++) {
        zo_w g = expert_slice(&L->gate_exps, idx[k], E), u = expert_slice(&L->up_exps, idx[k], E), d = expert_slice(&L->down_exps, idx[k], E);
        ffn_forward(m, &g, &u, &d, x, m->moe_out);
        for (int i = 0; i < m->hidden; i++) out[i] = out[i] + m->moe_out[i] * wk[k];
    }
}

/* Embedding row gather (+ scale): inference/arch_llama.go:246-342.       */
static int embed_row(zo_model *m, int token, float *out) {
    if (token < 0 || token >= m->vocab) return -1;
    int64_t rb = zo_row_bytes(m->embed.type, m->embed.cols);
    zo_dequant(m->embed.type, m->embed.data + (int64_t)token * rb, out, m->embed.cols);
    if (m->embed_scale > 0) for (int i = 0; i < m->hidden; i++) out[i] *= m->embed_scale;
    return 0;
}

/* One token through the whole stack at position m->pos; fills m->logits
 * when want_logits.  Prefill is the same arithmetic applied token by token
 * (rows are independent everywhere except attention, which is causal).  */
ZO_API int zo_model_forward(zo_model *m, int token, int want_logits) {
    if (m->pos >= m->max_seq) return -2;
    if (embed_row(m, token, m->hid)) return -1;
    int hd = m->hd, half = hd / 2, nq = m->n_q, nkv = m->n_kv;
    int qdim = nq * hd, kvdim = nkv * hd, rep = nq / nkv;
    float scale = (float)(1.0 / sqrt((double)hd));          /* sdpa.go:227 */
    int pos = m->pos;
    for (int li = 0; li < m->layers; li++) {
        zo_layer *L = &m->L[li];
        /* 1. input RMSNorm (arch_common.go:158-164) */
        zo_rmsnorm(m->normed, m->hid, (const float *)L->attn_norm.data, m->hidden, m->eps);
        /* 2. q/k/v projections (grouped_query_attention.go:524-549) */
        float *q = m->qkv, *k = m->qkv + qdim, *v = k + kvdim;
        wgemv(&L->q, m->normed, q);
        wgemv(&L->k, m->normed, k);
        wgemv(&L->v, m->normed, v);
        const float *cs = L->cos_tbl + (size_t)pos * half, *sn = L->sin_tbl + (size_t)pos * half;
        /* per-head q/k RMSNorm (Gemma 3) then half-split RoPE
         * (grouped_query_attention.go:579-827, fused_qk_norm_rope.cu:16-79) */
        for (int h = 0; h < nq + nkv; h++) {
            float *x = (h < nq) ? q + h * hd : k + (h - nq) * hd;
            if (m->qk_norm) {
                const float *w = (const float *)((h < nq) ? L->q_norm.data : L->k_norm.data);
                zo_rmsnorm(m->rowbuf, x, w, hd, m->eps);
                zo_rope(x, m->rowbuf, cs, sn, half, hd);
            } else {
                memcpy(m->rowbuf, x, (size_t)hd * 4);
                zo_rope(x, m->rowbuf, cs, sn, half, hd);
            }
        }
        /* cache.Update (tensor_cache.go:205-262); the fp16 cache converts on the write (offset_memcpy_fp16,
         * tensor_cache.go:224-238) and every later read sees the rounded values */
        if (m->kv_f16) {
            for (int i = 0; i < kvdim; i++) { k[i] = (float)(_Float16)k[i]; v[i] = (float)(_Float16)v[i]; }
        }
        memcpy(L->kc + (size_t)pos * kvdim, k, (size_t)kvdim * 4);
        memcpy(L->vc + (size_t)pos * kvdim, v, (size_t)kvdim * 4);
        /* A decode step (seqLen == 1) attends the whole cache, no sliding window; the prompt pass (seqLen > 1) of a
         * sliding-window model adds BuildCausalSlidingWindowMask: row i sees j <= i with i - j < window
         * (grouped_query_attention.go:1070-1081,1395-1415).  The additive -1e9 underflows to an exact 0 weight in f32, so
         * the masked rows are simply left out here. */
        int lo = 0;
        if (m->in_prefill && m->prefill_window > 0 && pos + 1 > m->prefill_window) lo = pos + 1 - m->prefill_window;
        for (int h = 0; h < nq; h++) {
            int kvh = h / rep;
            zo_attn_decode_head(m->attn + h * hd, q + h * hd, L->kc + (size_t)lo * kvdim + kvh * hd, L->vc + (size_t)lo * kvdim + kvh * hd,
                                pos + 1 - lo, hd, kvdim, scale, m->scores);
        }
        wgemv(&L->o, m->attn, m->proj);
        /* 3. Gemma 3 post-attention norm (arch_common.go:404-416) */
        if (m->post_norm) {
            zo_rmsnorm(m->rowbuf, m->proj, (const float *)L->post_attn_norm.data, m->hidden, m->eps);
            memcpy(m->proj, m->rowbuf, (size_t)m->hidden * 4);
        }
        /* 4. fused add + pre-FFN norm (fused_add_rmsnorm_node.go:32-47) */
        zo_add_rmsnorm(m->normed, m->res, m->proj, m->hid, (const float *)L->ffn_norm.data, m->hidden, m->eps);
        /* 5. FFN / MoE */
        if (L->router.type >= 0) moe_forward(m, L, m->normed, m->proj);
        else ffn_forward(m, &L->gate, &L->up, &L->down, m->normed, m->proj);
        /* 6. post-FFN norm+add (Gemma 3) or residual add */
        if (m->post_norm) zo_norm_add(m->hid, m->proj, (const float *)L->post_ffw_norm.data, m->res, m->hidden, m->eps);
        else for (int i = 0; i < m->hidden; i++) m->hid[i] = m->proj[i] + m->res[i];
    }
    m->pos++;
    if (want_logits) {
        zo_rmsnorm(m->normed, m->hid, (const float *)m->out_norm.data, m->hidden, m->eps);
        wgemv(&m->lm_head, m->normed, m->logits);
        if (m->softcap > 0) zo_softcap(m->logits, m->vocab, m->softcap);
    }
    return 0;
}

/* Greedy generation following generate/session.go:84-268: reset, prefill
 * the prompt, sample, then n_new-1 decode steps.  Returns tokens written. */
ZO_API void zo_model_set_kv_f16(zo_model *m, int on) { m->kv_f16 = on ? 1 : 0; }

/* The prompt pass: the reference runs the whole prompt through one Forward (seqLen = n), which is where a sliding-window
 * model applies its mask (grouped_query_attention.go:1074-1077).  Token by token here, with the same visibility. */
ZO_API int zo_model_prefill(zo_model *m, const int *prompt, int n_prompt, int want_logits) {
    m->in_prefill = 1;
    int rc = 0;
    for (int i = 0; i < n_prompt && !rc; i++) rc = zo_model_forward(m, prompt[i], want_logits && i == n_prompt - 1);
    m->in_prefill = 0;
    return rc;
}

ZO_API int zo_model_generate(zo_model *m, const int *prompt, int n_prompt, int n_new, int *out_tokens) {
    zo_model_reset(m);
    if (zo_model_prefill(m, prompt, n_prompt, 1)) return -1;
    int produced = 0;
    int tok = zo_argmax(m->logits, m->vocab);
    out_tokens[produced++] = tok;
    while (produced < n_new) {
        if (zo_model_forward(m, tok, 1)) break;
        tok = zo_argmax(m->logits, m->vocab);
        out_tokens[produced++] = tok;
    }
    return produced;
}

/* The harness (bench.py --impl reference under torchrun, which exports OMP_NUM_THREADS=1) sets the team size itself. */
ZO_API void zo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ZO_API int zo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ZO_API void zo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
