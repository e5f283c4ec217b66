// Element-wise, bit-exact dequantisation of raw GGUF blocks (device side).
// Every product below is exact in f32 (fp16 scale x small integer), so the
// result does not depend on evaluation order or FMA contraction; the single
// rounding is the final subtraction of the Q4_K/Q5_K min term.
// Formats: model/gguf/loader.go:140-190; gemv_q4k.cu:8-21; gemv_q5k.cu:7-23;
// gemv_q6k.cu:7-25; Q4_0 nibble order internal/xblas/q4dot.go:10-29.
#pragma once
#include "zb_common.cuh"

namespace zb {

enum { kF32 = 0, kF16 = 1, kQ4_0 = 2, kQ8_0 = 8, kQ4_K = 12, kQ5_K = 13, kQ6_K = 14 };

__host__ __device__ inline int block_elems(int t) {
    return (t == kF32 || t == kF16) ? 1 : ((t == kQ4_0 || t == kQ8_0) ? 32 : ((t == kQ4_K || t == kQ5_K || t == kQ6_K) ? 256 : 0));
}
__host__ __device__ inline int block_bytes(int t) {
    switch (t) {
        case kF32: return 4;
        case kF16: return 2;
        case kQ4_0: return 18;
        case kQ8_0: return 34;
        case kQ4_K: return 144;
        case kQ5_K: return 176;
        case kQ6_K: return 210;
    }
    return 0;
}

__device__ __forceinline__ uint16_t ld_u16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// fp16 -> f32 exactly as the reference CPU engine decodes block scales
// (internal/xblas/q4dot.go:53-80): -0 decodes to +0 and Inf/NaN decode to 0.
// Used on the bit-exact dequantisation path; the GEMV kernels use the plain
// hardware conversion (those encodings contribute 0 to a dot product either way,
// and never occur in a valid GGUF).
__device__ __forceinline__ float h2f_ref(uint16_t bits) {
    if ((bits & 0x7FFFu) == 0 || (bits & 0x7C00u) == 0x7C00u) return 0.0f;
    return h2f(bits);
}

__device__ __forceinline__ void kq_unpack(const uint8_t* sc, int j, int& s, int& m) {
    if (j < 4) {
        s = sc[j] & 63;
        m = sc[4 + j] & 63;
    } else {
        s = (sc[4 + j] & 0xF) | ((sc[j - 4] >> 6) << 4);
        m = (sc[4 + j] >> 4) | ((sc[j] >> 6) << 4);
    }
}

// Element i (0-based within the tensor) of a raw GGUF tensor of type t.
__device__ inline float deq_raw(int t, const uint8_t* base, int64_t i) {
    switch (t) {
        case kF32: return reinterpret_cast<const float*>(base)[i];
        case kF16: return h2f_ref(ld_u16(base + 2 * i));
        case kQ4_0: {
            const uint8_t* b = base + (i >> 5) * 18;
            int j = (int)(i & 31);
            uint8_t q = b[2 + (j & 15)];
            int v = (j < 16 ? (q & 0xF) : (q >> 4)) - 8;
            return __fmul_rn((float)v, h2f_ref(ld_u16(b)));
        }
        case kQ8_0: {
            const uint8_t* b = base + (i >> 5) * 34;
            return __fmul_rn((float)(int8_t)b[2 + (i & 31)], h2f_ref(ld_u16(b)));
        }
        case kQ4_K:
        case kQ5_K: {
            const uint8_t* b = base + (i >> 8) * (t == kQ4_K ? 144 : 176);
            int e = (int)(i & 255), g = e >> 6, w = e & 63, l = w & 31, hi = w >> 5;
            int s, m;
            kq_unpack(b + 4, 2 * g + hi, s, m);
            uint8_t qb = b[16 + g * 32 + l];
            int q = hi ? (qb >> 4) : (qb & 0xF);
            if (t == kQ5_K) q |= ((b[144 + l] >> (2 * g + hi)) & 1) << 4;
            float d = h2f_ref(ld_u16(b)), dmin = h2f_ref(ld_u16(b + 2));
            return __fsub_rn(__fmul_rn(__fmul_rn(d, (float)s), (float)q), __fmul_rn(dmin, (float)m));
        }
        case kQ6_K: {
            const uint8_t* b = base + (i >> 8) * 210;
            int e = (int)(i & 255), half = e >> 7, w = e & 127, quarter = w >> 5, l = w & 31;
            uint8_t ql = b[half * 64 + (quarter & 1) * 32 + l];
            uint8_t qh = b[128 + half * 32 + l];
            int q = (int)(((quarter >> 1) ? (ql >> 4) : (ql & 0xF)) | (((qh >> (2 * quarter)) & 3) << 4)) - 32;
            int sc = (int)(int8_t)b[192 + half * 8 + quarter * 2 + (l >> 4)];
            return __fmul_rn(__fmul_rn(h2f_ref(ld_u16(b + 208)), (float)sc), (float)q);
        }
    }
    return 0.0f;
}

}  // namespace zb
