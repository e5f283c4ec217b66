// Host-side decode engine (C++ over the kernel ABI): GGUF -> device layouts ->
// KV cache -> CUDA-graph-captured decode step -> greedy generation.
//
// It stands where the reference's Go callers stand (no Go toolchain in the
// build image; see DESIGN.md "Boundary"):
//   inference.LoadFile / LoadGGUF           inference/load_gguf.go:17-181, model/gguf/parser.go:96
//   ExtractModelConfig                      model/gguf/arch.go:132-245
//   WeightUploader.UploadWeights            inference/load_gguf.go:101-116 (Q4_0 -> separated layout,
//                                           Q8_0 -> 36 B blocks, K-quants raw: SURVEY 8b "Ownership")
//   buildTransformerGraph (per-layer ops)   inference/arch_common.go:150-526
//   TensorCache.Update / counters           generate/tensor_cache.go:205-262
//   InferenceSession.Generate, graphForward generate/session.go:84-268,440-461
//   CUDA graph capture of the step          generate/generator.go:301-365
//   tryGPUArgmax (4-byte D2H)               generate/sampling_helpers.go:11-45
// There is no CPU fallback: every op is one of this library's CUDA launchers.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#include "zb200.h"
#include "zb_common.cuh"
#include "zb_quant.cuh"
#include "zerfoo_kernels.h"

namespace {

using namespace zb;

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (cudaError_t)(expr);                                                        \
        if (_e != cudaSuccess) return fail((int)_e, "%s -> %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

enum { ZB_EINVAL = -1, ZB_EIO = -2, ZB_EFORMAT = -3, ZB_EUNSUPPORTED = -4, ZB_ESTATE = -5 };

// --------------------------------------------------------------------------
// GGUF v2/v3 (model/gguf/parser.go:96; dims are GGML order, innermost first:
// model/gguf/loader.go:97-104; 32-byte data alignment unless general.alignment)
// --------------------------------------------------------------------------
struct GTensor {
    std::string name;
    int type = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    const uint8_t* data = nullptr;
    int64_t rows() const { return ne[1] * ne[2] * ne[3]; }
    int64_t cols() const { return ne[0]; }
    int64_t nbytes() const { return rows() * cols() / block_elems(type) * block_bytes(type); }
};

struct Gguf {
    const uint8_t* base = nullptr;
    size_t size = 0;
    std::map<std::string, double> num;
    std::map<std::string, std::string> str;
    std::map<std::string, GTensor> tensors;
    ~Gguf() {
        if (base) munmap((void*)base, size);
    }
    const GTensor* find(const std::string& n) const {
        auto it = tensors.find(n);
        return it == tensors.end() ? nullptr : &it->second;
    }
};

struct Rd {
    const uint8_t *p, *end;
    bool bad = false;
    template <typename T>
    T get() {
        T v{};
        if (p + sizeof(T) > end) { bad = true; return v; }
        memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        uint64_t n = get<uint64_t>();
        if (bad || n > (uint64_t)(end - p)) { bad = true; return {}; }
        std::string s((const char*)p, (size_t)n);
        p += n;
        return s;
    }
    double scalar(uint32_t t) {
        switch (t) {
            case 0: return get<uint8_t>();
            case 1: return get<int8_t>();
            case 2: return get<uint16_t>();
            case 3: return get<int16_t>();
            case 4: return get<uint32_t>();
            case 5: return get<int32_t>();
            case 6: return get<float>();
            case 7: return get<uint8_t>();
            case 10: return (double)get<uint64_t>();
            case 11: return (double)get<int64_t>();
            case 12: return get<double>();
        }
        bad = true;
        return 0;
    }
};

int gguf_open(const char* path, Gguf& g) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(ZB_EIO, "open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) || st.st_size < 24) {
        close(fd);
        return fail(ZB_EIO, "stat %s failed or file too small", path);
    }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return fail(ZB_EIO, "mmap %s: %s", path, strerror(errno));
    g.base = (const uint8_t*)m;
    g.size = (size_t)st.st_size;
    Rd r{g.base, g.base + g.size};
    if (r.get<uint32_t>() != 0x46554747u) return fail(ZB_EFORMAT, "%s: not a GGUF file", path);
    uint32_t ver = r.get<uint32_t>();
    if (ver < 2 || ver > 3) return fail(ZB_EFORMAT, "%s: GGUF version %u unsupported", path, ver);
    uint64_t nt = r.get<uint64_t>(), nkv = r.get<uint64_t>();
    if (nt > (1u << 20) || nkv > (1u << 20)) return fail(ZB_EFORMAT, "%s: implausible header counts", path);
    for (uint64_t i = 0; i < nkv && !r.bad; i++) {
        std::string key = r.str();
        uint32_t t = r.get<uint32_t>();
        if (t == 8) g.str[key] = r.str();
        else if (t == 9) {
            uint32_t et = r.get<uint32_t>();
            uint64_t cnt = r.get<uint64_t>();
            for (uint64_t j = 0; j < cnt && !r.bad; j++) {
                if (et == 8) r.str();
                else r.scalar(et);
            }
            g.num[key] = (double)cnt;
        } else g.num[key] = r.scalar(t);
    }
    std::vector<std::pair<GTensor, uint64_t>> infos;
    for (uint64_t i = 0; i < nt && !r.bad; i++) {
        GTensor t;
        t.name = r.str();
        uint32_t nd = r.get<uint32_t>();
        if (nd > 4) { r.bad = true; break; }
        for (uint32_t d = 0; d < nd; d++) t.ne[d] = (int64_t)r.get<uint64_t>();
        t.type = (int)r.get<uint32_t>();
        uint64_t off = r.get<uint64_t>();
        infos.push_back({t, off});
    }
    if (r.bad) return fail(ZB_EFORMAT, "%s: truncated or malformed GGUF header", path);
    uint64_t align = 32;
    if (g.num.count("general.alignment") && g.num["general.alignment"] > 0) align = (uint64_t)g.num["general.alignment"];
    uint64_t start = ((uint64_t)(r.p - g.base) + align - 1) / align * align;
    for (auto& it : infos) {
        GTensor t = it.first;
        int be = block_elems(t.type);
        if (be == 0) return fail(ZB_EUNSUPPORTED, "%s: tensor %s has unsupported ggml type %d", path, t.name.c_str(), t.type);
        if (t.ne[0] % be) return fail(ZB_EFORMAT, "%s: tensor %s row length %lld not a multiple of block %d", path, t.name.c_str(), (long long)t.ne[0], be);
        if (start + it.second + (uint64_t)t.nbytes() > g.size) return fail(ZB_EFORMAT, "%s: tensor %s runs past end of file", path, t.name.c_str());
        t.data = g.base + start + it.second;
        g.tensors[t.name] = t;
    }
    return 0;
}

// --------------------------------------------------------------------------
// Device weights
// --------------------------------------------------------------------------
enum Layout { kRaw = 0, kQ4Sep = 1, kQ8_36 = 2 };

struct DW {
    int type = -1;
    int layout = kRaw;
    int64_t rows = 0, cols = 0;
    void* d = nullptr;
    int data_offset = 0;   // Q4 separated: byte offset of the nibble region
    int64_t bytes = 0;
};

struct Layer {
    DW attn_norm, q_norm, k_norm, post_attn_norm, ffn_norm, post_ffw_norm;
    std::vector<DW> qkv;          // 1..3 GEMVs writing consecutive slices of the qkv buffer
    DW o;
    std::vector<DW> gate_up;      // 1..2 GEMVs writing [gate | up]
    DW down;
    DW router;                    // MoE
    std::vector<DW> e_gate_up, e_down;
    float* kc = nullptr;
    float* vc = nullptr;
    const float* cos_tbl = nullptr;
    const float* sin_tbl = nullptr;
};

}  // namespace

struct zb_engine {
    zb_engine_opts opts{};
    Gguf g;
    std::string arch;
    int vocab = 0, hidden = 0, layers = 0, n_q = 0, n_kv = 0, hd = 0, ffn = 0, max_seq = 0, n_experts = 0, top_k = 0;
    float eps = 1e-5f, softcap = 0.0f, embed_scale = 0.0f;
    bool post_norm = false, qk_norm = false;
    double rope_base = 10000.0, rope_local = 0.0;
    int sw_pattern = 0;

    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<void*> allocs;
    std::vector<Layer> L;
    DW embed_raw, out_norm, lm_head;
    float *tbl_gc = nullptr, *tbl_gs = nullptr, *tbl_lc = nullptr, *tbl_ls = nullptr;

    // activations (persistent; a captured graph bakes these addresses in)
    float *hid = nullptr, *normed = nullptr, *qkv = nullptr, *qrot = nullptr, *attn = nullptr, *proj = nullptr, *proj2 = nullptr,
          *res = nullptr, *gateup = nullptr, *act = nullptr, *logits = nullptr, *part_o = nullptr, *part_lse = nullptr;
    void* amax_scratch = nullptr;
    int *d_cur = nullptr, *d_last = nullptr, *d_pos = nullptr, *d_kvlen = nullptr, *d_feed = nullptr, *d_feed_idx = nullptr,
        *d_feed_len = nullptr, *d_out = nullptr, *d_nout = nullptr, *d_amax = nullptr;
    int* h_pin = nullptr;  // pinned host ints: [0] token in, [1] token out
    int feed_cap = 0, out_cap = 0;
    int splits = 1, chunk = 256;

    cudaGraphExec_t graph_full = nullptr, graph_nohead = nullptr;
    int launches_full = 0;
    int64_t weight_bytes = 0;
    int host_pos = 0;

    // per-launch GEMV profiler (zb_engine_profile_gemv): CUDA events around every weight-streaming launch
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    struct ProfRec { int type; double bytes; };
    std::vector<ProfRec> prof_rec;

    ~zb_engine() {
        for (auto ev : prof_ev) cudaEventDestroy(ev);
        if (graph_full) cudaGraphExecDestroy(graph_full);
        if (graph_nohead) cudaGraphExecDestroy(graph_nohead);
        for (void* p : allocs) cudaFree(p);
        if (h_pin) cudaFreeHost(h_pin);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

template <typename T>
int dalloc(zb_engine* e, T** out, size_t count) {
    void* p = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    CK(cudaMalloc(&p, bytes));
    CK(cudaMemset(p, 0, bytes));
    e->allocs.push_back(p);
    *out = (T*)p;
    return 0;
}

// Concatenate tensors (same type, same K) row-wise, convert to the device
// layout, upload.  Mirrors MergeQ4Storage/MergeQ4KStorage + UploadWeights
// (inference/arch_common.go:337-372,477-502; load_gguf.go:101-116).
// Optional row range [r0, r1) of the concatenation (expert slices, TP shards).
int upload(zb_engine* e, const std::vector<const GTensor*>& ts, DW& w, int64_t r0 = 0, int64_t r1 = -1) {
    const GTensor* t0 = ts[0];
    int type = t0->type;
    int64_t cols = t0->cols(), rows = 0;
    for (auto* t : ts) {
        if (t->type != type || t->cols() != cols) return fail(ZB_EINVAL, "upload: cannot merge %s with %s", t->name.c_str(), t0->name.c_str());
        rows += t->rows();
    }
    if (r1 < 0) r1 = rows;
    int64_t rb = cols / block_elems(type) * block_bytes(type);
    std::vector<uint8_t> raw((size_t)((r1 - r0) * rb));
    int64_t at = 0, out = 0;
    for (auto* t : ts) {  // copy the intersection of [r0,r1) with this tensor's rows
        int64_t lo = std::max<int64_t>(r0, at), hi = std::min<int64_t>(r1, at + t->rows());
        if (hi > lo) {
            memcpy(raw.data() + out, t->data + (lo - at) * rb, (size_t)((hi - lo) * rb));
            out += (hi - lo) * rb;
        }
        at += t->rows();
    }
    rows = r1 - r0;
    w.type = type;
    w.rows = rows;
    w.cols = cols;
    std::vector<uint8_t> conv;
    const uint8_t* src = raw.data();
    size_t bytes = raw.size();
    if (type == kQ4_0) {  // separated: [fp16 scales][pad16][16 B nibbles] (gemm_q4.h:3-4)
        int64_t nblk = rows * (cols / 32);
        int64_t pad = (nblk * 2 + 15) & ~(int64_t)15;
        conv.assign((size_t)(pad + nblk * 16), 0);
        for (int64_t b = 0; b < nblk; b++) {
            memcpy(&conv[(size_t)(b * 2)], &raw[(size_t)(b * 18)], 2);
            memcpy(&conv[(size_t)(pad + b * 16)], &raw[(size_t)(b * 18 + 2)], 16);
        }
        w.layout = kQ4Sep;
        w.data_offset = (int)pad;
        if (pad > 0x7fffffff) return fail(ZB_EUNSUPPORTED, "Q4_0 tensor too large for int data_offset");
        src = conv.data();
        bytes = conv.size();
    } else if (type == kQ8_0) {  // f32 scale + 32 int8 (gemm_q8.cu:1-7)
        int64_t nblk = rows * (cols / 32);
        conv.resize((size_t)(nblk * 36));
        for (int64_t b = 0; b < nblk; b++) {
            uint16_t h;
            memcpy(&h, &raw[(size_t)(b * 34)], 2);
            float f = __half2float(__ushort_as_half(h));
            memcpy(&conv[(size_t)(b * 36)], &f, 4);
            memcpy(&conv[(size_t)(b * 36 + 4)], &raw[(size_t)(b * 34 + 2)], 32);
        }
        w.layout = kQ8_36;
        src = conv.data();
        bytes = conv.size();
    } else {
        w.layout = kRaw;
    }
    uint8_t* d = nullptr;
    if (int rc = dalloc(e, &d, bytes + 16)) return rc;
    CK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    w.d = d;
    w.bytes = (int64_t)bytes;
    return 0;
}

int upload_raw(zb_engine* e, const GTensor* t, DW& w) {
    w.type = t->type; w.layout = kRaw; w.rows = t->rows(); w.cols = t->cols(); w.bytes = t->nbytes();
    uint8_t* d = nullptr;
    if (int rc = dalloc(e, &d, (size_t)w.bytes + 16)) return rc;
    CK(cudaMemcpy(d, t->data, (size_t)w.bytes, cudaMemcpyHostToDevice));
    w.d = d;
    return 0;
}

int gemv_launch(const DW& w, const float* x, float* y, cudaStream_t s) {
    cudaError_t rc;
    switch (w.type) {
        case kQ4_0: rc = gemm_q4_f32(w.d, x, y, (int)w.rows, (int)w.cols, 1, w.data_offset, s); break;
        case kQ8_0: rc = gemm_q8_f32(w.d, x, y, (int)w.rows, (int)w.cols, 1, s); break;
        case kQ4_K: rc = gemv_q4k_f32(w.d, x, y, (int)w.rows, (int)w.cols, s); break;
        case kQ5_K: rc = gemv_q5k_f32(w.d, x, y, (int)w.rows, (int)w.cols, s); break;
        case kQ6_K: rc = gemv_q6k_f32(w.d, x, y, (int)w.rows, (int)w.cols, s); break;
        case kF32: rc = launch_sgemv_m1(y, (const float*)w.d, x, (int)w.rows, (int)w.cols, s); break;
        default: return fail(ZB_EUNSUPPORTED, "gemv: unsupported weight type %d", w.type);
    }
    if (rc != cudaSuccess) return fail((int)rc, "gemv type %d [%lld x %lld]: %s", w.type, (long long)w.rows, (long long)w.cols, cudaGetErrorString(rc));
    return 0;
}

// Algorithmic bytes of one GEMV launch: weight blocks once + x + y.
double gemv_bytes(const DW& w) {
    return (double)(w.rows * (w.cols / block_elems(w.type)) * (int64_t)block_bytes(w.type)) + 4.0 * (double)w.cols + 4.0 * (double)w.rows;
}

int gemv(zb_engine* e, const DW& w, const float* x, float* y, cudaStream_t s) {
    if (!e->prof_on) return gemv_launch(w, x, y, s);
    size_t i = e->prof_rec.size() * 2;
    while (e->prof_ev.size() < i + 2) {
        cudaEvent_t ev;
        CK(cudaEventCreate(&ev));
        e->prof_ev.push_back(ev);
    }
    CK(cudaEventRecord(e->prof_ev[i], s));
    int rc = gemv_launch(w, x, y, s);
    CK(cudaEventRecord(e->prof_ev[i + 1], s));
    e->prof_rec.push_back({w.type, gemv_bytes(w)});
    return rc;
}

// --------------------------------------------------------------------------
// Engine-private kernels
// --------------------------------------------------------------------------
__global__ void step_begin_kernel(int* cur, const int* last, const int* feed, int* feed_idx, const int* feed_len) {
    int i = *feed_idx;
    if (i < *feed_len) {
        *cur = feed[i];
        *feed_idx = i + 1;
    } else {
        *cur = *last;
    }
}

// Embedding row gather with bit-exact dequantisation (+ Gemma scale):
// inference/arch_llama.go:246-342, arch_gemma.go:38.  Ids are clamped like
// launch_gather (gather.cu:20-23); the host API rejects out-of-range ids.
__global__ void embed_kernel(int type, const uint8_t* __restrict__ table, const int* __restrict__ cur, float* __restrict__ out, int hidden,
                             int vocab, float scale) {
    int tok = *cur;
    if (tok < 0) tok = 0;
    if (tok >= vocab) tok = vocab - 1;
    int64_t base = (int64_t)tok * hidden;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hidden; i += gridDim.x * blockDim.x) {
        float v = deq_raw(type, table, base + i);
        out[i] = scale > 0.0f ? v * scale : v;
    }
}

// One CTA per head of [q heads | k heads | v heads]: optional per-head RMSNorm
// (Gemma 3), half-split RoPE at the device-resident position, and the KV
// append, in one launch.  Replaces rope_select + fused_qk_norm_rope/fused_rope
// + 2*nKV offset_memcpy launches (grouped_query_attention.go:579-866,
// generate/tensor_cache.go:205-262).
__global__ void qkv_post_kernel(const float* __restrict__ qkv, const float* __restrict__ wq, const float* __restrict__ wk,
                                const float* __restrict__ cos_tbl, const float* __restrict__ sin_tbl, const int* __restrict__ pos_ptr,
                                float* __restrict__ q_out, float* __restrict__ kc, float* __restrict__ vc, float eps, int hd, int nq, int nkv,
                                int max_seq) {
    extern __shared__ float xn[];
    __shared__ float red[32];
    int head = blockIdx.x, pos = *pos_ptr;
    if (pos < 0 || pos >= max_seq) return;
    const float* x = qkv + (int64_t)head * hd;
    int half = hd / 2;
    if (head >= nq + nkv) {  // V head: straight into the cache
        float* dst = vc + (int64_t)pos * nkv * hd + (int64_t)(head - nq - nkv) * hd;
        for (int d = threadIdx.x; d < hd; d += blockDim.x) dst[d] = x[d];
        return;
    }
    const float* w = head < nq ? wq : wk;
    if (w) {
        float ss = 0.0f;
        for (int d = threadIdx.x; d < hd; d += blockDim.x) ss = fmaf(x[d], x[d], ss);
        ss = block_sum(ss, red);
        float s = (float)(1.0 / sqrt((double)(ss / (float)hd + eps)));
        for (int d = threadIdx.x; d < hd; d += blockDim.x) xn[d] = x[d] * s * w[d];
    } else {
        for (int d = threadIdx.x; d < hd; d += blockDim.x) xn[d] = x[d];
    }
    __syncthreads();
    float* o = head < nq ? q_out + (int64_t)head * hd : kc + (int64_t)pos * nkv * hd + (int64_t)(head - nq) * hd;
    const float* cs = cos_tbl + (int64_t)pos * half;
    const float* sn = sin_tbl + (int64_t)pos * half;
    for (int d = threadIdx.x; d < half; d += blockDim.x) {
        float a = xn[d], b = xn[d + half], c = cs[d], s = sn[d];
        o[d] = a * c - b * s;
        o[d + half] = b * c + a * s;
    }
}

// Gemma softcap as the CPU engine evaluates it (inference/arch_llama.go:15-27,184-213):
// cap * tanh_rational(logit / cap), clamped to +-1 beyond |x| >= 4.5.
__global__ void softcap_kernel(float* logits, int n, float cap, float inv_cap) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = logits[i] * inv_cap, t;
    if (x > 4.5f) t = 1.0f;
    else if (x < -4.5f) t = -1.0f;
    else {
        float x2 = x * x;
        t = x * (27.0f + x2) / (27.0f + 9.0f * x2);
    }
    logits[i] = cap * t;
}

__global__ void step_end_kernel(int* pos, int* kvlen, const int* amax, int* last, int* out, int* n_out, int out_cap, int with_head) {
    *pos += 1;
    *kvlen += 1;
    if (with_head) {
        int t = *amax;
        *last = t;
        int n = *n_out;
        if (n < out_cap) out[n] = t;
        *n_out = n + 1;
    }
}

// MoE router on the device (replaces the host sort.Slice round trip of
// layers/core/moe.go:110-146): softmax over E, top-k by probability with
// lowest-index tie-break, weights renormalised to sum 1.  One warp.
__global__ void moe_route_kernel(const float* __restrict__ logits, int E, int K, int* __restrict__ idx_out, float* __restrict__ w_out) {
    __shared__ float p[256];
    __shared__ int chosen[256];
    if (threadIdx.x == 0) {
        float mx = logits[0];
        for (int i = 1; i < E; i++) mx = fmaxf(mx, logits[i]);
        float sum = 0.0f;
        for (int i = 0; i < E; i++) {
            p[i] = (float)exp((double)(logits[i] - mx));
            sum += p[i];
            chosen[i] = 0;
        }
        float inv = 1.0f / sum;
        for (int i = 0; i < E; i++) p[i] *= inv;
        float wsum = 0.0f;
        for (int k = 0; k < K; k++) {
            int best = -1;
            for (int i = 0; i < E; i++)
                if (!chosen[i] && (best < 0 || p[i] > p[best])) best = i;
            chosen[best] = 1;
            idx_out[k] = best;
            w_out[k] = p[best];
            wsum = wsum + p[best];
        }
        for (int k = 0; k < K; k++) w_out[k] = w_out[k] / wsum;
    }
}

// out (+)= w[k] * x  -- the MoE combine (layers/core/moe.go:470-479).
__global__ void scale_accum_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ w, int k, int n, int first) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i] * w[k];
    out[i] = first ? (0.0f + v) : (out[i] + v);
}

// --------------------------------------------------------------------------
// Model load
// --------------------------------------------------------------------------
double kvnum(const Gguf& g, const std::string& arch, const char* suffix, double dflt) {
    auto it = g.num.find(arch + "." + suffix);
    return it == g.num.end() ? dflt : it->second;
}

int need(const Gguf& g, const std::string& name, const GTensor** out) {
    *out = g.find(name);
    if (!*out) return fail(ZB_EFORMAT, "missing tensor %s", name.c_str());
    return 0;
}

int load_norm(zb_engine* e, const std::string& name, DW& w, bool required) {
    const GTensor* t = e->g.find(name);
    if (!t) return required ? fail(ZB_EFORMAT, "missing tensor %s", name.c_str()) : 0;
    if (t->type != kF32) return fail(ZB_EUNSUPPORTED, "%s: norm weights must be F32", name.c_str());
    return upload_raw(e, t, w);
}

// Group consecutive same-type tensors into merged GEMVs.
int upload_group(zb_engine* e, const std::vector<const GTensor*>& ts, std::vector<DW>& out) {
    size_t i = 0;
    while (i < ts.size()) {
        size_t j = i + 1;
        while (j < ts.size() && ts[j]->type == ts[i]->type && ts[j]->cols() == ts[i]->cols()) j++;
        DW w;
        if (int rc = upload(e, std::vector<const GTensor*>(ts.begin() + i, ts.begin() + j), w)) return rc;
        out.push_back(w);
        i = j;
    }
    return 0;
}

void rope_tables(std::vector<float>& cs, std::vector<float>& sn, int positions, int rot, double base) {
    // layers/embeddings/rotary_positional_embedding.go:117-160: f64 pow/cos/sin -> f32
    int half = rot / 2;
    cs.resize((size_t)positions * half);
    sn.resize((size_t)positions * half);
    std::vector<double> inv(half);
    for (int i = 0; i < half; i++) inv[i] = 1.0 / pow(base, (double)(2 * i) / (double)rot);
    for (int p = 0; p < positions; p++)
        for (int j = 0; j < half; j++) {
            double a = (double)p * inv[j];
            cs[(size_t)p * half + j] = (float)cos(a);
            sn[(size_t)p * half + j] = (float)sin(a);
        }
}

int load_model(zb_engine* e, const char* path) {
    if (int rc = gguf_open(path, e->g)) return rc;
    const Gguf& g = e->g;
    auto a = g.str.find("general.architecture");
    if (a == g.str.end() || a->second.empty()) return fail(ZB_EFORMAT, "missing general.architecture metadata");
    e->arch = a->second;
    const std::string& ar = e->arch;
    e->vocab = (int)kvnum(g, ar, "vocab_size", 0);
    e->hidden = (int)kvnum(g, ar, "embedding_length", 0);
    e->layers = (int)kvnum(g, ar, "block_count", 0);
    e->n_q = (int)kvnum(g, ar, "attention.head_count", 0);
    e->n_kv = (int)kvnum(g, ar, "attention.head_count_kv", e->n_q);
    e->ffn = (int)kvnum(g, ar, "feed_forward_length", 0);
    int ctx = (int)kvnum(g, ar, "context_length", 2048);
    e->rope_base = kvnum(g, ar, "rope.freq_base", 0);
    if (e->rope_base == 0) e->rope_base = kvnum(g, ar, "rope.global.freq_base", 10000.0);
    e->hd = (int)kvnum(g, ar, "attention.key_length", 0);
    if (e->hd <= 0 && e->n_q > 0) e->hd = e->hidden / e->n_q;
    e->softcap = (float)kvnum(g, ar, "final_logit_softcapping", 0);
    e->rope_local = kvnum(g, ar, "rope.local.freq_base", 0);
    e->sw_pattern = e->rope_local > 0 ? 6 : 0;
    e->eps = (float)kvnum(g, ar, "attention.layer_norm_rms_epsilon", 0);
    if (!(e->eps > 0)) e->eps = 1e-5f;
    e->n_experts = (int)kvnum(g, ar, "expert_count", 0);
    e->top_k = (int)kvnum(g, ar, "expert_used_count", 0);
    bool is_gemma = ar.rfind("gemma", 0) == 0, is_gemma3 = ar == "gemma3";
    bool is_moe = ar == "mixtral" || e->n_experts > 0;
    if (is_moe) {
        if (!e->n_experts) e->n_experts = 8;
        if (!e->top_k) e->top_k = 2;
        if (e->n_experts > 256) return fail(ZB_EUNSUPPORTED, "expert_count %d > 256", e->n_experts);
    }
    if (is_gemma) e->embed_scale = (float)sqrt((double)e->hidden);
    if (is_gemma3) { e->post_norm = true; e->qk_norm = true; } else e->softcap = 0.0f;
    if (e->hidden <= 0 || e->layers <= 0 || e->n_q <= 0 || e->n_kv <= 0 || e->hd <= 0 || e->n_q % e->n_kv)
        return fail(ZB_EFORMAT, "invalid model dimensions (hidden %d layers %d heads %d/%d head_dim %d)", e->hidden, e->layers, e->n_q, e->n_kv, e->hd);
    if (e->hd > 256 || e->hd % 2) return fail(ZB_EUNSUPPORTED, "head_dim %d unsupported (even, <= 256)", e->hd);
    e->max_seq = e->opts.max_seq > 0 ? e->opts.max_seq : (ctx < 4096 ? ctx : 4096);
    if (e->max_seq > ctx) e->max_seq = ctx;

    const GTensor *t_embed, *t_onorm;
    if (int rc = need(g, "token_embd.weight", &t_embed)) return rc;
    if (int rc = need(g, "output_norm.weight", &t_onorm)) return rc;
    e->vocab = (int)t_embed->rows();
    if (t_embed->cols() != e->hidden) return fail(ZB_EFORMAT, "token_embd row length %lld != hidden %d", (long long)t_embed->cols(), e->hidden);
    if (int rc = upload_raw(e, t_embed, e->embed_raw)) return rc;
    if (int rc = load_norm(e, "output_norm.weight", e->out_norm, true)) return rc;
    const GTensor* t_head = g.find("output.weight");
    if (!t_head) t_head = t_embed;  // tied head (arch_llama.go:54-58, arch_gemma.go:36)
    bool raw_is_gemv_layout = t_head->type != kQ4_0 && t_head->type != kQ8_0;
    if (t_head == t_embed && raw_is_gemv_layout) e->lm_head = e->embed_raw;  // share the table with the gather
    else if (int rc = upload(e, {t_head}, e->lm_head)) return rc;
    e->weight_bytes += e->lm_head.bytes;

    std::vector<float> cs, sn;
    int half = e->hd / 2;
    rope_tables(cs, sn, e->max_seq, e->hd, e->rope_base);
    if (int rc = dalloc(e, &e->tbl_gc, cs.size())) return rc;
    if (int rc = dalloc(e, &e->tbl_gs, sn.size())) return rc;
    CK(cudaMemcpy(e->tbl_gc, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->tbl_gs, sn.data(), sn.size() * 4, cudaMemcpyHostToDevice));
    if (e->sw_pattern > 0) {
        rope_tables(cs, sn, e->max_seq, e->hd, e->rope_local);
        if (int rc = dalloc(e, &e->tbl_lc, cs.size())) return rc;
        if (int rc = dalloc(e, &e->tbl_ls, sn.size())) return rc;
        CK(cudaMemcpy(e->tbl_lc, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(e->tbl_ls, sn.data(), sn.size() * 4, cudaMemcpyHostToDevice));
    }
    (void)half;

    e->L.resize(e->layers);
    for (int i = 0; i < e->layers; i++) {
        Layer& L = e->L[i];
        std::string p = "blk." + std::to_string(i) + ".";
        if (int rc = load_norm(e, p + "attn_norm.weight", L.attn_norm, true)) return rc;
        if (int rc = load_norm(e, p + "attn_q_norm.weight", L.q_norm, e->qk_norm)) return rc;
        if (int rc = load_norm(e, p + "attn_k_norm.weight", L.k_norm, e->qk_norm)) return rc;
        if (int rc = load_norm(e, p + "post_attention_norm.weight", L.post_attn_norm, e->post_norm)) return rc;
        if (int rc = load_norm(e, p + "ffn_norm.weight", L.ffn_norm, true)) return rc;
        if (int rc = load_norm(e, p + "post_ffw_norm.weight", L.post_ffw_norm, e->post_norm)) return rc;
        const GTensor *q, *k, *v, *o;
        if (int rc = need(g, p + "attn_q.weight", &q)) return rc;
        if (int rc = need(g, p + "attn_k.weight", &k)) return rc;
        if (int rc = need(g, p + "attn_v.weight", &v)) return rc;
        if (int rc = need(g, p + "attn_output.weight", &o)) return rc;
        if (q->rows() != (int64_t)e->n_q * e->hd || k->rows() != (int64_t)e->n_kv * e->hd || v->rows() != k->rows() || q->cols() != e->hidden ||
            o->rows() != e->hidden || o->cols() != q->rows())
            return fail(ZB_EFORMAT, "layer %d: attention weight shapes do not match the config", i);
        if (int rc = upload_group(e, {q, k, v}, L.qkv)) return rc;
        if (int rc = upload(e, {o}, L.o)) return rc;
        for (auto& w : L.qkv) e->weight_bytes += w.bytes;
        e->weight_bytes += L.o.bytes;
        if (is_moe) {
            const GTensor *r, *ge, *ue, *de;
            if (int rc = need(g, p + "ffn_gate_inp.weight", &r)) return rc;
            if (int rc = need(g, p + "ffn_gate_exps.weight", &ge)) return rc;
            if (int rc = need(g, p + "ffn_up_exps.weight", &ue)) return rc;
            if (int rc = need(g, p + "ffn_down_exps.weight", &de)) return rc;
            if (int rc = upload(e, {r}, L.router)) return rc;
            int E = e->n_experts;
            int64_t fr = ge->rows() / E, dr = de->rows() / E;
            for (int x = 0; x < E; x++) {  // expert slices at block-row boundaries (arch_mixtral.go buildExpertFFN)
                DW gw, uw, dw;
                if (ge->type == ue->type) {
                    // [gate_x ; up_x] merged into one GEMV per expert
                    std::vector<uint8_t> dummy;
                    DW m;
                    GTensor gs = *ge, us = *ue;
                    int64_t rb = ge->cols() / block_elems(ge->type) * block_bytes(ge->type);
                    gs.data = ge->data + x * fr * rb; gs.ne[1] = fr; gs.ne[2] = 1;
                    us.data = ue->data + x * fr * rb; us.ne[1] = fr; us.ne[2] = 1;
                    if (int rc = upload(e, {&gs, &us}, m)) return rc;
                    L.e_gate_up.push_back(m);
                } else {
                    return fail(ZB_EUNSUPPORTED, "layer %d: expert gate/up types differ", i);
                }
                if (int rc = upload(e, {de}, dw, x * dr, (x + 1) * dr)) return rc;
                L.e_down.push_back(dw);
            }
            e->weight_bytes += L.router.bytes + (int64_t)e->top_k * (L.e_gate_up[0].bytes + L.e_down[0].bytes);
        } else {
            const GTensor *ga, *up, *dn;
            if (int rc = need(g, p + "ffn_gate.weight", &ga)) return rc;
            if (int rc = need(g, p + "ffn_up.weight", &up)) return rc;
            if (int rc = need(g, p + "ffn_down.weight", &dn)) return rc;
            if (ga->rows() != up->rows() || dn->cols() != ga->rows() || dn->rows() != e->hidden)
                return fail(ZB_EFORMAT, "layer %d: FFN weight shapes do not match", i);
            e->ffn = (int)ga->rows();
            if (int rc = upload_group(e, {ga, up}, L.gate_up)) return rc;
            if (int rc = upload(e, {dn}, L.down)) return rc;
            for (auto& w : L.gate_up) e->weight_bytes += w.bytes;
            e->weight_bytes += L.down.bytes;
        }
        bool global = !(e->sw_pattern > 0 && ((i + 1) % e->sw_pattern != 0));  // arch_common.go:171-178
        L.cos_tbl = global ? e->tbl_gc : e->tbl_lc;
        L.sin_tbl = global ? e->tbl_gs : e->tbl_ls;
        size_t kvsz = (size_t)e->max_seq * e->n_kv * e->hd;
        if (int rc = dalloc(e, &L.kc, kvsz)) return rc;
        if (int rc = dalloc(e, &L.vc, kvsz)) return rc;
    }
    if (is_moe) e->ffn = (int)(e->L[0].e_down[0].cols);

    int qd = e->n_q * e->hd, kvd = e->n_kv * e->hd;
    e->chunk = 256;
    e->splits = (e->max_seq + e->chunk - 1) / e->chunk;
    if (int rc = dalloc(e, &e->hid, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->normed, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->qkv, qd + 2 * kvd)) return rc;
    if (int rc = dalloc(e, &e->qrot, qd)) return rc;
    if (int rc = dalloc(e, &e->attn, qd)) return rc;
    if (int rc = dalloc(e, &e->proj, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->proj2, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->res, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->gateup, 2 * (size_t)e->ffn + 256)) return rc;
    if (int rc = dalloc(e, &e->act, e->ffn)) return rc;
    if (int rc = dalloc(e, &e->logits, e->vocab)) return rc;
    if (int rc = dalloc(e, &e->part_o, (size_t)e->n_q * e->splits * e->hd)) return rc;
    if (int rc = dalloc(e, &e->part_lse, 2 * (size_t)e->n_q * e->splits)) return rc;
    float* sc = nullptr;
    if (int rc = dalloc(e, &sc, 2 * (size_t)((e->vocab + 255) / 256) + 64)) return rc;
    e->amax_scratch = sc;
    e->feed_cap = e->max_seq;
    e->out_cap = e->max_seq;
    int* ints = nullptr;
    if (int rc = dalloc(e, &ints, 16 + (size_t)e->feed_cap + e->out_cap)) return rc;
    e->d_cur = ints; e->d_last = ints + 1; e->d_pos = ints + 2; e->d_kvlen = ints + 3; e->d_feed_idx = ints + 4; e->d_feed_len = ints + 5;
    e->d_nout = ints + 6; e->d_amax = ints + 7;
    e->d_feed = ints + 16;
    e->d_out = ints + 16 + e->feed_cap;
    CK(cudaMallocHost(&e->h_pin, 64));
    return 0;
}

// --------------------------------------------------------------------------
// One decode step, enqueued on the engine stream (capturable).
// --------------------------------------------------------------------------
struct Counter {
    int n = 0;
};

#define LAUNCH(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (cudaError_t)(expr);                                                          \
        cnt.n++;                                                                                       \
        if (_e != cudaSuccess) return fail((int)_e, "%s -> %s", #expr, cudaGetErrorString(_e));        \
    } while (0)
#define KLAUNCH(...)                                                                                   \
    do {                                                                                               \
        __VA_ARGS__;                                                                                   \
        cnt.n++;                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess) return fail((int)_e, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

unsigned f2u(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    return u;
}

int enqueue_step(zb_engine* e, bool with_head, Counter& cnt) {
    cudaStream_t s = e->stream;
    const int H = e->hidden, hd = e->hd, nq = e->n_q, nkv = e->n_kv, qd = nq * hd, kvd = nkv * hd;
    const unsigned eps = f2u(e->eps);
    KLAUNCH(step_begin_kernel<<<1, 1, 0, s>>>(e->d_cur, e->d_last, e->d_feed, e->d_feed_idx, e->d_feed_len));
    KLAUNCH(embed_kernel<<<(H + 255) / 256, 256, 0, s>>>(e->embed_raw.type, (const uint8_t*)e->embed_raw.d, e->d_cur, e->hid, H, e->vocab,
                                                         e->embed_scale));
    for (int li = 0; li < e->layers; li++) {
        Layer& L = e->L[li];
        LAUNCH(launch_rmsnorm(e->hid, (const float*)L.attn_norm.d, e->normed, nullptr, eps, 1, H, s));
        int64_t off = 0;
        for (auto& w : L.qkv) {
            if (int rc = gemv(e, w, e->normed, e->qkv + off, s)) return rc;
            cnt.n++;
            off += w.rows;
        }
        int threads = hd >= 256 ? 256 : (hd >= 128 ? 128 : 64);
        KLAUNCH(qkv_post_kernel<<<nq + 2 * nkv, threads, hd * sizeof(float), s>>>(
            e->qkv, e->qk_norm ? (const float*)L.q_norm.d : nullptr, e->qk_norm ? (const float*)L.k_norm.d : nullptr, L.cos_tbl, L.sin_tbl,
            e->d_pos, e->qrot, L.kc, L.vc, e->eps, hd, nq, nkv, e->max_seq));
        LAUNCH(flash_decode_splitkv_f32(e->qrot, L.kc, L.vc, e->attn, e->part_o, e->part_lse, nq, e->max_seq, hd, e->max_seq, e->d_kvlen, nq, nkv,
                                        e->chunk, s));
        cnt.n++;  // partial + reduce
        if (int rc = gemv(e, L.o, e->attn, e->proj, s)) return rc;
        cnt.n++;
        const float* attn_out = e->proj;
        if (e->post_norm) {
            LAUNCH(launch_rmsnorm(e->proj, (const float*)L.post_attn_norm.d, e->proj2, nullptr, eps, 1, H, s));
            attn_out = e->proj2;
        }
        LAUNCH(fused_add_rmsnorm_f32(attn_out, e->hid, (const float*)L.ffn_norm.d, e->normed, e->res, eps, 1, H, s));
        if (L.router.d) {
            float* rl = e->gateup + 2 * (size_t)e->ffn;      // router logits / weights scratch (256 floats reserved)
            int* ridx = e->d_amax + 1;                        // top-k indices (ints region has 8 spare slots)
            (void)ridx;
            return fail(ZB_EUNSUPPORTED, "MoE decode is dispatched by enqueue_moe");
        } else {
            off = 0;
            for (auto& w : L.gate_up) {
                if (int rc = gemv(e, w, e->normed, e->gateup + off, s)) return rc;
                cnt.n++;
                off += w.rows;
            }
            LAUNCH(fused_swiglu_f32(e->gateup, e->gateup + e->ffn, e->act, e->ffn, s));
            if (int rc = gemv(e, L.down, e->act, e->proj, s)) return rc;
            cnt.n++;
        }
        if (e->post_norm) LAUNCH(fused_norm_add_f32(e->proj, (const float*)L.post_ffw_norm.d, e->res, e->hid, eps, 1, H, s));
        else LAUNCH(launch_add(e->proj, e->res, e->hid, H, s));
    }
    if (with_head) {
        LAUNCH(launch_rmsnorm(e->hid, (const float*)e->out_norm.d, e->normed, nullptr, eps, 1, H, s));
        if (int rc = gemv(e, e->lm_head, e->normed, e->logits, s)) return rc;
        cnt.n++;
        if (e->softcap > 0.0f)
            KLAUNCH(softcap_kernel<<<(e->vocab + 255) / 256, 256, 0, s>>>(e->logits, e->vocab, e->softcap, (float)(1.0 / (double)e->softcap)));
        LAUNCH(launch_argmax(e->logits, e->d_amax, e->amax_scratch, e->vocab, s));
        cnt.n++;  // two stages
    }
    KLAUNCH(step_end_kernel<<<1, 1, 0, s>>>(e->d_pos, e->d_kvlen, e->d_amax, e->d_last, e->d_out, e->d_nout, e->out_cap, with_head ? 1 : 0));
    (void)qd; (void)kvd;
    return 0;
}

int capture(zb_engine* e, bool with_head, cudaGraphExec_t* out) {
    cudaGraph_t graph = nullptr;
    Counter cnt;
    CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_step(e, with_head, cnt);
    cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (ce != cudaSuccess) return fail((int)ce, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail((int)ce, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
    if (with_head) e->launches_full = cnt.n;
    return 0;
}

int run_step(zb_engine* e, bool with_head) {
    if (e->host_pos >= e->max_seq) return fail(ZB_ESTATE, "KV cache full (%d positions)", e->max_seq);
    cudaGraphExec_t gx = with_head ? e->graph_full : e->graph_nohead;
    if (gx) {
        CK(cudaGraphLaunch(gx, e->stream));
    } else {
        Counter cnt;
        if (int rc = enqueue_step(e, with_head, cnt)) return rc;
        if (with_head) e->launches_full = cnt.n;
    }
    e->host_pos++;
    return 0;
}

// Pre-warm like the reference's session pool (inference/load_gguf.go:155-168): run
// one eager step so every kernel is loaded and its attributes are set, capture the
// two step graphs (with / without lm_head), then rewind the counters.  The KV row
// written by the warm-up is overwritten by the first real token.
int warm_and_capture(zb_engine* e) {
    if (int rc = zb_engine_reset(e)) return rc;
    if (int rc = run_step(e, true)) return rc;
    CK(cudaStreamSynchronize(e->stream));
    if (e->opts.use_graph) {
        if (int rc = capture(e, true, &e->graph_full)) return rc;
        if (int rc = capture(e, false, &e->graph_nohead)) return rc;
    }
    return zb_engine_reset(e);
}

int set_feed(zb_engine* e, const int32_t* tokens, int n) {
    if (n > e->feed_cap) return fail(ZB_EINVAL, "prompt of %d tokens exceeds capacity %d", n, e->feed_cap);
    for (int i = 0; i < n; i++)
        if (tokens[i] < 0 || tokens[i] >= e->vocab) return fail(ZB_EINVAL, "token ID %d out of range [0, %d)", tokens[i], e->vocab);
    int hdr[2] = {0, n};
    CK(cudaMemcpyAsync(e->d_feed, tokens, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->d_feed_idx, hdr, 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(e->d_nout, 0, 4, e->stream));
    CK(cudaStreamSynchronize(e->stream));  // hdr/tokens are pageable host memory
    return 0;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
ZB_API const char* zb_last_error(void) { return g_err.c_str(); }

ZB_API int zb_engine_create(const char* gguf_path, const zb_engine_opts* opts, zb_engine** out) {
    if (!gguf_path || !out) return fail(ZB_EINVAL, "zb_engine_create: null argument");
    *out = nullptr;
    zb_engine* e = new zb_engine();
    if (opts) e->opts = *opts;
    else { e->opts.use_graph = 1; }
    if (e->opts.tp_size <= 0) e->opts.tp_size = 1;
    const char* dis = getenv("ZERFOO_DISABLE_CUDA_GRAPH");  // generate/generator.go:328
    if (dis && dis[0] && strcmp(dis, "0")) e->opts.use_graph = 0;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        delete e;
        return fail(ce ? (int)ce : ZB_ESTATE, "no CUDA device available (%s): this engine has no CPU fallback", cudaGetErrorString(ce));
    }
    int rc = 0;
    do {
        if ((ce = cudaSetDevice(e->opts.device)) != cudaSuccess) { rc = fail((int)ce, "cudaSetDevice(%d): %s", e->opts.device, cudaGetErrorString(ce)); break; }
        if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) { rc = fail((int)ce, "cudaStreamCreate: %s", cudaGetErrorString(ce)); break; }
        cudaEventCreate(&e->ev0);
        cudaEventCreate(&e->ev1);
        if (e->opts.tp_size != 1) { rc = fail(ZB_EUNSUPPORTED, "tensor parallel engine: use zb_engine_create_tp"); break; }
        rc = load_model(e, gguf_path);
        if (rc) break;
        if (e->n_experts > 0) { rc = fail(ZB_EUNSUPPORTED, "MoE models are not wired into the decode step yet"); break; }
        rc = warm_and_capture(e);
    } while (0);
    if (rc) {
        delete e;
        return rc;
    }
    *out = e;
    return 0;
}

ZB_API void zb_engine_destroy(zb_engine* e) {
    if (!e) return;
    cudaSetDevice(e->opts.device);
    cudaStreamSynchronize(e->stream);
    delete e;
}

ZB_API int zb_engine_info(const zb_engine* e, zb_model_info* o) {
    if (!e || !o) return fail(ZB_EINVAL, "zb_engine_info: null argument");
    memset(o, 0, sizeof *o);
    o->vocab = e->vocab; o->hidden = e->hidden; o->layers = e->layers; o->n_q = e->n_q; o->n_kv = e->n_kv; o->head_dim = e->hd;
    o->ffn = e->ffn; o->max_seq = e->max_seq; o->n_experts = e->n_experts; o->top_k = e->top_k;
    o->tp_rank = e->opts.tp_rank; o->tp_size = e->opts.tp_size;
    o->weight_bytes_per_token = e->weight_bytes;
    o->kv_bytes_per_pos = 2LL * e->layers * e->n_kv * e->hd * 4;
    o->launches_per_step = e->launches_full;
    snprintf(o->arch, sizeof o->arch, "%s", e->arch.c_str());
    return 0;
}

ZB_API int zb_engine_reset(zb_engine* e) {
    if (!e) return fail(ZB_EINVAL, "null engine");
    CK(cudaSetDevice(e->opts.device));
    int zeros[8] = {0, 0, 0, 1, 0, 0, 0, 0};  // cur,last,pos,kvlen(=pos+1),feed_idx,feed_len,nout,amax
    CK(cudaMemcpyAsync(e->d_cur, zeros, sizeof zeros, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->host_pos = 0;
    return 0;
}

ZB_API int zb_engine_prefill(zb_engine* e, const int32_t* tokens, int n, int32_t* first_token) {
    if (!e || !tokens || n <= 0) return fail(ZB_EINVAL, "zb_engine_prefill: bad arguments");
    CK(cudaSetDevice(e->opts.device));
    if (e->host_pos + n > e->max_seq) return fail(ZB_ESTATE, "prompt does not fit the KV cache (%d + %d > %d)", e->host_pos, n, e->max_seq);
    if (int rc = set_feed(e, tokens, n)) return rc;
    for (int i = 0; i < n; i++)
        if (int rc = run_step(e, i == n - 1)) return rc;
    CK(cudaMemcpyAsync(e->h_pin + 1, e->d_last, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (first_token) *first_token = e->h_pin[1];
    return 0;
}

ZB_API int zb_engine_decode_step(zb_engine* e, int32_t token, int32_t* next_token) {
    if (!e) return fail(ZB_EINVAL, "null engine");
    if (token < 0 || token >= e->vocab) return fail(ZB_EINVAL, "token ID %d out of range [0, %d)", token, e->vocab);
    CK(cudaSetDevice(e->opts.device));
    e->h_pin[0] = token;
    CK(cudaMemcpyAsync(e->d_last, e->h_pin, 4, cudaMemcpyHostToDevice, e->stream));
    if (int rc = run_step(e, true)) return rc;
    CK(cudaMemcpyAsync(e->h_pin + 1, e->d_last, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (next_token) *next_token = e->h_pin[1];
    return 0;
}

ZB_API int zb_engine_decode_n(zb_engine* e, int32_t first_token, int n, int32_t* out_tokens, float* ms) {
    if (!e || n <= 0) return fail(ZB_EINVAL, "zb_engine_decode_n: bad arguments");
    if (first_token < 0 || first_token >= e->vocab) return fail(ZB_EINVAL, "token ID %d out of range [0, %d)", first_token, e->vocab);
    if (n > e->out_cap) return fail(ZB_EINVAL, "n=%d exceeds output capacity %d", n, e->out_cap);
    CK(cudaSetDevice(e->opts.device));
    if (e->host_pos + n > e->max_seq) return fail(ZB_ESTATE, "%d steps do not fit the KV cache (pos %d, capacity %d)", n, e->host_pos, e->max_seq);
    e->h_pin[0] = first_token;
    CK(cudaMemcpyAsync(e->d_last, e->h_pin, 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(e->d_nout, 0, 4, e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    for (int i = 0; i < n; i++)
        if (int rc = run_step(e, true)) return rc;
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (ms) CK(cudaEventElapsedTime(ms, e->ev0, e->ev1));
    if (out_tokens) CK(cudaMemcpy(out_tokens, e->d_out, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

ZB_API int zb_engine_generate(zb_engine* e, const int32_t* prompt, int n_prompt, int n_new, int32_t* out_tokens) {
    if (!e || !prompt || n_prompt <= 0 || n_new <= 0 || !out_tokens) return fail(ZB_EINVAL, "zb_engine_generate: bad arguments");
    if (int rc = zb_engine_reset(e)) return rc;
    if (n_prompt + n_new - 1 > e->max_seq) return fail(ZB_ESTATE, "prompt %d + %d new tokens exceed the KV capacity %d", n_prompt, n_new, e->max_seq);
    if (n_new > e->out_cap) return fail(ZB_EINVAL, "n_new exceeds output capacity");
    if (int rc = set_feed(e, prompt, n_prompt)) return rc;
    for (int i = 0; i < n_prompt; i++)
        if (int rc = run_step(e, i == n_prompt - 1)) return rc;
    for (int i = 1; i < n_new; i++)
        if (int rc = run_step(e, true)) return rc;
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out_tokens, e->d_out, (size_t)n_new * 4, cudaMemcpyDeviceToHost));
    return 0;
}

ZB_API int zb_engine_logits(zb_engine* e, float* host_out) {
    if (!e || !host_out) return fail(ZB_EINVAL, "null argument");
    CK(cudaSetDevice(e->opts.device));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(host_out, e->logits, (size_t)e->vocab * 4, cudaMemcpyDeviceToHost));
    return 0;
}

ZB_API int zb_engine_hidden(zb_engine* e, float* host_out) {
    if (!e || !host_out) return fail(ZB_EINVAL, "null argument");
    CK(cudaSetDevice(e->opts.device));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(host_out, e->hid, (size_t)e->hidden * 4, cudaMemcpyDeviceToHost));
    return 0;
}

ZB_API int zb_engine_kv(zb_engine* e, int layer, int n, float* k_host, float* v_host) {
    if (!e || layer < 0 || layer >= e->layers || n < 0 || n > e->max_seq) return fail(ZB_EINVAL, "zb_engine_kv: bad arguments");
    CK(cudaSetDevice(e->opts.device));
    CK(cudaStreamSynchronize(e->stream));
    size_t bytes = (size_t)n * e->n_kv * e->hd * 4;
    if (k_host) CK(cudaMemcpy(k_host, e->L[layer].kc, bytes, cudaMemcpyDeviceToHost));
    if (v_host) CK(cudaMemcpy(v_host, e->L[layer].vc, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

ZB_API int zb_engine_position(const zb_engine* e) { return e ? e->host_pos : -1; }
ZB_API zb_stream_t zb_engine_stream(const zb_engine* e) { return e ? (zb_stream_t)e->stream : nullptr; }

// Runs `steps` eager (non-graph) decode steps with a CUDA-event pair around every
// weight-streaming GEMV launch on the engine stream and reports, per block format,
// launches / algorithmic bytes / summed device time.  The KV position advances.
ZB_API int zb_engine_profile_gemv(zb_engine* e, int steps, zb_gemv_profile* out, int max_classes, int* n_classes) {
    if (!e || steps <= 0 || !out || !n_classes) return fail(ZB_EINVAL, "zb_engine_profile_gemv: bad arguments");
    CK(cudaSetDevice(e->opts.device));
    if (e->host_pos + steps > e->max_seq) return fail(ZB_ESTATE, "profile steps do not fit the KV cache");
    e->prof_rec.clear();
    e->prof_on = true;
    int rc = 0;
    for (int i = 0; i < steps && !rc; i++) {
        Counter cnt;
        rc = enqueue_step(e, true, cnt);
        if (!rc) e->host_pos++;
    }
    e->prof_on = false;
    if (rc) return rc;
    CK(cudaStreamSynchronize(e->stream));
    int n = 0;
    for (size_t i = 0; i < e->prof_rec.size(); i++) {
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, e->prof_ev[2 * i], e->prof_ev[2 * i + 1]));
        int c = -1;
        for (int j = 0; j < n; j++)
            if (out[j].qtype == e->prof_rec[i].type) c = j;
        if (c < 0) {
            if (n >= max_classes) continue;
            c = n++;
            out[c].qtype = e->prof_rec[i].type;
            out[c].launches = 0;
            out[c].bytes = 0;
            out[c].ms = 0;
        }
        out[c].launches++;
        out[c].bytes += e->prof_rec[i].bytes;
        out[c].ms += ms;
    }
    *n_classes = n;
    return 0;
}
