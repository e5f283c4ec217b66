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
#include <dlfcn.h>
#include <unistd.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "zb200.h"
#include "zb_common.cuh"
#include "zb_quant.cuh"
#include "zb_stream.cuh"
#include "zb_mega.cuh"
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
        for (uint32_t d = 0; d < nd; d++) {
            const uint64_t ne = r.get<uint64_t>();
            if (ne == 0 || ne > (uint64_t)INT32_MAX) { r.bad = true; break; }   // untrusted file: dimensions must fit the int arithmetic below
            t.ne[d] = (int64_t)ne;
        }
        if (r.bad) break;
        t.type = (int)r.get<uint32_t>();
        uint64_t off = r.get<uint64_t>();
        infos.push_back({t, off});
    }
    if (r.bad) return fail(ZB_EFORMAT, "%s: truncated or malformed GGUF header", path);
    uint64_t align = 32;
    if (g.num.count("general.alignment") && g.num["general.alignment"] > 0) align = (uint64_t)g.num["general.alignment"];
    if (align == 0 || align > 4096 || (align & (align - 1))) return fail(ZB_EFORMAT, "%s: general.alignment %llu is not a power of two <= 4096", path, (unsigned long long)align);
    uint64_t start = ((uint64_t)(r.p - g.base) + align - 1) / align * align;
    if (start > g.size) return fail(ZB_EFORMAT, "%s: header runs past end of file", path);
    for (auto& it : infos) {
        GTensor t = it.first;
        int be = block_elems(t.type);
        if (be == 0) return fail(ZB_EUNSUPPORTED, "%s: tensor %s has unsupported ggml type %d", path, t.name.c_str(), t.type);
        if (t.ne[0] % be) return fail(ZB_EFORMAT, "%s: tensor %s row length %lld not a multiple of block %d", path, t.name.c_str(), (long long)t.ne[0], be);
        {   // checked sizes: a crafted header must not wrap the arithmetic into passing
            unsigned long long elems = 1, nbytes = 0;
            bool ovf = false;
            for (int d = 0; d < 4; d++) ovf = ovf || __builtin_mul_overflow(elems, (unsigned long long)(t.ne[d] > 0 ? t.ne[d] : 1), &elems);
            ovf = ovf || __builtin_mul_overflow(elems / (unsigned long long)be, (unsigned long long)block_bytes(t.type), &nbytes);
            const uint64_t avail = g.size - start;
            if (ovf || elems / (unsigned long long)t.ne[0] > (unsigned long long)INT32_MAX || it.second > avail || nbytes > avail - it.second ||
                (uint64_t)t.nbytes() != nbytes)
                return fail(ZB_EFORMAT, "%s: tensor %s runs past end of file", path, t.name.c_str());
        }
        t.data = g.base + start + it.second;
        g.tensors[t.name] = t;
    }
    return 0;
}

// --------------------------------------------------------------------------
// NCCL, resolved at run time like the reference does (distributed/nccl.go:60-98 dlopens libnccl and binds
// ncclCommInitRank / ncclAllReduce ...).  Note the x86-64 hazard SURVEY 2.3 records: ncclUniqueId is a 128-byte
// struct passed BY VALUE; calling through a typed C++ function pointer gets the SysV ABI right.
// --------------------------------------------------------------------------
struct NcclId { char internal[128]; };
struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;

int nccl_load() {
    if (g_nccl.lib) return 0;
    const char* names[] = {getenv("ZB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
        if (n && n[0] && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return fail(ZB_EIO, "cannot dlopen libnccl.so.2 (set ZB_NCCL_LIB): %s", dlerror());
    g_nccl.lib = h;
    *(void**)&g_nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void**)&g_nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)&g_nccl.AllReduce = dlsym(h, "ncclAllReduce");
    *(void**)&g_nccl.AllGather = dlsym(h, "ncclAllGather");
    *(void**)&g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void**)&g_nccl.GetErrorString = dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.AllGather)
        return fail(ZB_EIO, "libnccl is missing ncclGetUniqueId/ncclCommInitRank/ncclAllReduce/ncclAllGather");
    return 0;
}
#define NCCLK(expr)                                                                                                 \
    do {                                                                                                            \
        int _r = (expr);                                                                                            \
        if (_r != 0) return fail(ZB_EIO, "%s -> nccl error %d (%s)", #expr, _r, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
    } while (0)

// --------------------------------------------------------------------------
// Tensor-parallel shards (inference/parallel/tensor_parallel.go:40-48): column-parallel layers split the OUTPUT
// dimension = whole block rows of the GGUF matrix; row-parallel layers split the INPUT dimension = a block-aligned
// column range of every row.  Pure host byte slicing, shared with the CPU tests through zb_tp_shard_host.
// --------------------------------------------------------------------------
// out must hold (r1-r0) * ((c1-c0)/block_elems*block_bytes) bytes
int shard_bytes(int type, const uint8_t* raw, int64_t rows, int64_t cols, int64_t r0, int64_t r1, int64_t c0, int64_t c1, uint8_t* out) {
    const int be = block_elems(type), bb = block_bytes(type);
    if (be == 0 || r0 < 0 || r1 > rows || r0 > r1 || c0 < 0 || c1 > cols || c0 > c1 || c0 % be || c1 % be || cols % be) return -1;
    const int64_t rb = cols / be * bb, ob = (c1 - c0) / be * bb, co = c0 / be * bb;
    for (int64_t r = r0; r < r1; r++) memcpy(out + (size_t)((r - r0) * ob), raw + (size_t)(r * rb + co), (size_t)ob);
    return 0;
}

struct Slicer {  // keeps the sliced copies alive while load_model runs
    std::vector<std::vector<uint8_t>*> bufs;
    std::vector<GTensor*> ts;
    ~Slicer() {
        for (auto* b : bufs) delete b;
        for (auto* t : ts) delete t;
    }
    const GTensor* make(const GTensor* t, int64_t r0, int64_t r1, int64_t c0, int64_t c1) {
        auto* buf = new std::vector<uint8_t>((size_t)((r1 - r0) * row_bytes_of(t->type, c1 - c0)));
        bufs.push_back(buf);
        if (shard_bytes(t->type, t->data, t->rows(), t->cols(), r0, r1, c0, c1, buf->data())) return nullptr;
        GTensor* n = new GTensor(*t);
        ts.push_back(n);
        n->data = buf->data();
        n->ne[0] = c1 - c0; n->ne[1] = r1 - r0; n->ne[2] = 1; n->ne[3] = 1;
        return n;
    }
    static int64_t row_bytes_of(int type, int64_t cols) { return cols / block_elems(type) * block_bytes(type); }
    const GTensor* rows(const GTensor* t, int64_t r0, int64_t r1) { return make(t, r0, r1, 0, t->cols()); }
    const GTensor* cols(const GTensor* t, int64_t c0, int64_t c1) { return make(t, 0, t->rows(), c0, c1); }
};

// --------------------------------------------------------------------------
// Device weights: every quantized matrix lives in the stream layout
// (zb_stream.cuh): 16-B aligned block rows + separate fp16 block scales.
// --------------------------------------------------------------------------
struct DW {
    int type = -1;
    int64_t rows = 0, cols = 0;
    uint8_t* main = nullptr;   // stream layout (quantized) ...
    uint8_t* aux = nullptr;
    uint8_t* mma = nullptr;    // ... and, for the batch-1 tensor-core GEMV, as 16-row block-tiles (gemv_mma.cu)
    void* d = nullptr;         // ... or plain F32 (norm gains, router) / raw GGUF blocks (embedding gather)
    int64_t bytes = 0;         // GGUF bytes of the matrix (the algorithmic traffic of one GEMV)
    int64_t e_main_stride = 0, e_aux_stride = 0;  // MoE expert stack
    int64_t e_mma_stride = 0;                     // bytes between experts in the block-tile copy
    bool pairs = false;        // rows interleaved (gate_0, up_0, gate_1, up_1, ...): the GEMV epilogue applies SwiGLU
    int kslabs = 0;            // > 1: a matrix too wide for one tensor-core GEMV (K > 16384), stored as this many column slabs
                               // [rows x cols] stacked like an expert stack; slab s multiplies x[s*cols .. (s+1)*cols), the partial
                               // outputs are summed by the consumer's prologue (mix_n = kslabs, weights 1)
};

struct Layer {
    DW attn_norm, q_norm, k_norm, post_attn_norm, ffn_norm, post_ffw_norm;
    std::vector<DW> qkv;          // 1..3 GEMVs writing consecutive slices of the qkv buffer
    DW o;
    std::vector<DW> gate_up;      // 1..2 GEMVs writing [gate | up]
    DW down;
    DW router;                    // MoE: F32 [E, H]
    DW e_gate_up, e_down;         // MoE: expert stacks ([gate_x ; up_x] merged per expert)
    float* kc = nullptr;          // [n_kv][max_seq][hd]; f32, or fp16 bytes behind the same pointer when the engine's kv_f16 is set
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
    bool use_pdl = true;

    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<void*> allocs;
    std::vector<Layer> L;
    DW embed_raw, out_norm, lm_head;
    float *tbl_gc = nullptr, *tbl_gs = nullptr, *tbl_lc = nullptr, *tbl_ls = nullptr;

    // activations (persistent; a captured graph bakes these addresses in)
    float *hid = nullptr, *res = nullptr, *normed = nullptr, *qkv = nullptr, *attn = nullptr, *proj_o = nullptr, *proj = nullptr,
          *gateup = nullptr, *logits = nullptr, *part_o = nullptr, *part_ml = nullptr, *moe_y = nullptr, *rlogits = nullptr, *rw = nullptr;
    void* amax_scratch = nullptr;
    int *d_last = nullptr, *d_pos = nullptr, *d_feed = nullptr, *d_feed_idx = nullptr, *d_feed_len = nullptr, *d_out = nullptr,
        *d_nout = nullptr, *d_amax = nullptr, *d_ridx = nullptr, *d_ticket = nullptr;
    int* h_pin = nullptr;  // pinned host ints: [0] token in, [1] token out
    int feed_cap = 0, out_cap = 0;
    int chunk = 32, max_splits = 1;
    bool use_mma = true;             // batch-1 GEMVs of supported formats run on the tensor pipe (ZB_GEMV_TC=0: CUDA-core kernel only)
    int64_t mma_scratch_bytes = 0;
    void* mma_scratch = nullptr;     // row-tile tickets + split-tile partial sums, shared by all launches (stream-ordered)

    // persistent whole-token kernel (decode_mega.cu): dense single-GPU batch-1 models whose matrices all have block-tiles
    bool want_mega = false, mega = false;
    MegaCtl mctl{};
    int mega_ctas = 0;
    unsigned int* d_mega_bar = nullptr;   // [0] barrier arrivals, [1] launches since reset

    cudaGraphExec_t graph_full = nullptr, graph_nohead = nullptr;
    // short-context variant of the step: one long attention tile per KV head (16 warps, no split merge) while kv_len <= chunk_short
    cudaGraphExec_t graph_full_s = nullptr, graph_nohead_s = nullptr;
    bool kv_f16 = false;         // KV cache stored as fp16 (ZB_ENGINE_KV_F16 / ZB_KV_F16=1)
    int prefill_window = 0;      // sliding window of the prompt pass (0 = none); d_window_on = 1 while a prompt is being fed
    int* d_window_on = nullptr;
    int max_kslabs = 0;          // column-slab stacks (DW::kslabs): partial outputs [kslabs][hidden], identity slot table, unit weights
    float* slab_y = nullptr;
    int* d_iota = nullptr;
    float* d_fones = nullptr;
    int chunk_short = 0;
    bool attn_short = false;         // which variant enqueue_step builds
    int launches_full = 0;
    int64_t weight_bytes = 0;
    int host_pos = 0;
    const float* final_hid = nullptr;  // where the last step left the post-stack residual stream

    // ---- tensor parallel state (opts.tp_size > 1)
    int tp_rank = 0, tp_size = 1, n_q_global = 0, n_kv_global = 0, vocab_local = 0, experts_local = 0;
    void* nccl_comm = nullptr;
    float* logits_local = nullptr;
    // fused exchange over NVLink peer memory (cudaIpc): flags [2][P] then slots [2][P][xn] in one allocation per rank
    bool tp_fused = false;
    bool tp_push = false;        // one-shot all-reduce over peer memory (tp_allreduce_push_kernel) instead of ncclAllReduce
    int ar_site = 0;             // all-reduce sites enqueued so far in the step being built (2 per layer)
    int ar_sites_per_step = 0;   // = 2 * layers: every (step, site) pair gets its own epoch
    uint8_t* xchg_local = nullptr;
    uint8_t* xchg_peer[8] = {nullptr};
    size_t xchg_slots_off = 0;
    int xn = 0;                      // floats per slot
    float* d_ones = nullptr;
    int *d_step = nullptr, *d_xticket = nullptr;
    int* d_ridx_local = nullptr;

    // ---- batched decode state (opts.batch > 1)
    int B = 1, page = 16, max_blocks = 0, pool_blocks = 0, Bpad = 16;
    std::vector<int> h_btab, free_blocks, h_bpos;
    int *d_btab = nullptr, *d_bpos = nullptr, *d_btok = nullptr, *d_bamax = nullptr, *d_bout = nullptr, *d_bnout = nullptr;
    float *b_hid = nullptr, *b_res = nullptr, *b_qkv = nullptr, *b_attn = nullptr, *b_proj_o = nullptr, *b_proj = nullptr, *b_gateup = nullptr,
          *b_logits = nullptr;
    void *b_xhi = nullptr, *b_xlo = nullptr;
    int* h_bpin = nullptr;
    cudaGraphExec_t graph_batch = nullptr;
    int launches_batch = 0;
    void* pchunk = nullptr;   // ChunkBufs of the chunked prefill (allocated on first use)

    // per-launch GEMV profiler (zb_engine_profile_gemv): CUDA events around every weight-streaming launch
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    struct ProfRec { int type; double bytes; };
    std::vector<ProfRec> prof_rec;

    ~zb_engine() {
        for (auto ev : prof_ev) cudaEventDestroy(ev);
        if (graph_full) cudaGraphExecDestroy(graph_full);
        if (graph_nohead) cudaGraphExecDestroy(graph_nohead);
        if (graph_full_s) cudaGraphExecDestroy(graph_full_s);
        if (graph_nohead_s) cudaGraphExecDestroy(graph_nohead_s);
        if (graph_batch) cudaGraphExecDestroy(graph_batch);
        for (int i = 0; i < 8; i++)
            if (xchg_peer[i] && xchg_peer[i] != xchg_local) cudaIpcCloseMemHandle(xchg_peer[i]);
        if (nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(nccl_comm);
        if (h_bpin) cudaFreeHost(h_bpin);
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

int64_t row_bytes(int type, int64_t cols) { return cols / block_elems(type) * block_bytes(type); }

// Rows [r0, r1) of the row-wise concatenation of `ts` as raw GGUF blocks (host).
int gather_rows(const std::vector<const GTensor*>& ts, int64_t r0, int64_t r1, std::vector<uint8_t>& raw, int& type, int64_t& cols) {
    const GTensor* t0 = ts[0];
    type = t0->type;
    cols = t0->cols();
    int64_t rows = 0;
    for (auto* t : ts) {
        if (t->type != type || t->cols() != cols) return fail(ZB_EINVAL, "upload: cannot merge %s with %s", t->name.c_str(), t0->name.c_str());
        rows += t->rows();
    }
    if (r1 < 0) r1 = rows;
    int64_t rb = row_bytes(type, cols);
    raw.resize((size_t)((r1 - r0) * rb));
    int64_t at = 0, out = 0;
    for (auto* t : ts) {
        int64_t lo = std::max<int64_t>(r0, at), hi = std::min<int64_t>(r1, at + t->rows());
        if (hi > lo) {
            memcpy(raw.data() + out, t->data + (lo - at) * rb, (size_t)((hi - lo) * rb));
            out += (hi - lo) * rb;
        }
        at += t->rows();
    }
    return 0;
}

// Merge (MergeQ4Storage / MergeQ4KStorage, inference/arch_common.go:337-372,477-502), repack to the
// stream layout and upload (UploadWeights, load_gguf.go:101-116).  `experts` > 1: the rows are E equal
// expert slices that must stay addressable by index (arch_mixtral.go buildExpertFFN).
int upload_raw_rows(zb_engine* e, const std::vector<uint8_t>& raw, int type, int64_t rows, int64_t cols, DW& w, int experts = 1, int slots = 0) {
    if (int rc = zb_stream_check(type, (int)(rows / experts), (int)cols))
        return fail(ZB_EUNSUPPORTED, "matrix [%lld x %lld] of ggml type %d does not fit the streamed GEMV (rc %d)", (long long)rows, (long long)cols, type, rc);
    int64_t mb = 0, ab = 0;
    if (zb_stream_layout(type, (int)rows, (int)cols, &mb, &ab)) return fail(ZB_EUNSUPPORTED, "no stream layout for ggml type %d", type);
    std::vector<uint8_t> hm((size_t)mb), ha((size_t)ab, 0);
    if (zb_stream_repack_host(type, raw.data(), (int)rows, (int)cols, hm.data(), ab ? ha.data() : nullptr)) return fail(ZB_EUNSUPPORTED, "repack failed");
    w.type = type;
    w.rows = rows;
    w.cols = cols;
    w.bytes = (int64_t)raw.size();
    if (int rc = dalloc(e, &w.main, (size_t)mb + 64)) return rc;
    CK(cudaMemcpy(w.main, hm.data(), (size_t)mb, cudaMemcpyHostToDevice));
    if (ab) {
        if (int rc = dalloc(e, &w.aux, (size_t)ab + 64)) return rc;
        CK(cudaMemcpy(w.aux, ha.data(), (size_t)ab, cudaMemcpyHostToDevice));
    }
    // Tensor-core GEMV policy (measured, profiles/r01_gemv_mma_microbench.log): every K-quant matrix; Q4_0 only where it beats the
    // CUDA-core kernel -- long rows (K >= 4096) and streaming-size matrices (>= 32 MB, the tied lm_head); ZB_MMA_Q4_0_ALL=1 overrides.
    static const bool q40_all = getenv("ZB_MMA_Q4_0_ALL") && getenv("ZB_MMA_Q4_0_ALL")[0] == '1';
    // tensor-parallel shards take the same path (their fused-exchange launches -- xsite >= 0 in gemv() -- stay on the CUDA-core kernel)
    static const bool tp_mma = !(getenv("ZB_TP_MMA") && getenv("ZB_TP_MMA")[0] == '0');
    const bool mma_pays = type != kQ4_0 || q40_all || e->want_mega || cols >= 4096 || (int64_t)raw.size() >= (32ll << 20);
    // expert stacks: every expert must be a whole number of 16-row tiles; the launch runs top_k slots side by side
    const bool stack_ok = experts == 1 || ((rows / experts) % 16 == 0 && zb_mma_check(type, (int)(rows / experts), (int)cols) == 0);
    if (stack_ok && e->use_mma && mma_pays && (e->tp_size == 1 || tp_mma) && e->opts.batch <= 1 && zb_mma_check(type, (int)rows, (int)cols) == 0) {
        int64_t wb = 0, sb = 0;
        if (zb_mma_layout(type, (int)rows, (int)cols, &wb, &sb)) return fail(ZB_EUNSUPPORTED, "no block-tile layout for ggml type %d", type);
        if (experts > 1) {
            int64_t wb1 = 0;
            if (zb_mma_layout(type, (int)(rows / experts), (int)cols, &wb1, &sb)) return fail(ZB_EUNSUPPORTED, "no block-tile layout for an expert");
            if (wb1 * experts != wb) return fail(ZB_ESTATE, "expert block-tile stride mismatch");
            w.e_mma_stride = wb1;
            sb *= std::max(1, std::max(e->top_k, slots));
        }
        std::unique_ptr<uint8_t[]> ht(new uint8_t[(size_t)wb]);   // uninitialised: the repack writes every byte
        if (zb_mma_repack_host(type, raw.data(), (int)rows, (int)cols, ht.get())) return fail(ZB_EUNSUPPORTED, "block-tile repack failed");
        if (int rc = dalloc(e, &w.mma, (size_t)wb + 64)) return rc;
        CK(cudaMemcpy(w.mma, ht.get(), (size_t)wb, cudaMemcpyHostToDevice));
        if (sb > e->mma_scratch_bytes) e->mma_scratch_bytes = sb;
    }
    if (experts > 1) {
        int64_t er = rows / experts;
        w.rows = er;
        w.bytes = (int64_t)raw.size() / experts;
        w.e_main_stride = er * stream_main_bytes(type, (int)(cols / 32));
        w.e_aux_stride = er * stream_aux_bytes(type, (int)(cols / 32));
    }
    return 0;
}

// [gate ; up] -> rows (gate_0, up_0, gate_1, up_1, ...) so that one warp owns both halves of every SwiGLU pair
// (layers/core/ffn.go:190-201 merges gate and up into one MatMul; :234-239 applies GPUFusedSwiGLU to the halves).
void interleave_pairs(const uint8_t* gate, const uint8_t* up, int64_t rows, int64_t rb, uint8_t* out) {
    for (int64_t i = 0; i < rows; i++) {
        memcpy(out + (size_t)(2 * i) * rb, gate + (size_t)i * rb, (size_t)rb);
        memcpy(out + (size_t)(2 * i + 1) * rb, up + (size_t)i * rb, (size_t)rb);
    }
}

int upload(zb_engine* e, const std::vector<const GTensor*>& ts, DW& w, int64_t r0 = 0, int64_t r1 = -1) {
    std::vector<uint8_t> raw;
    int type;
    int64_t cols;
    if (int rc = gather_rows(ts, r0, r1, raw, type, cols)) return rc;
    return upload_raw_rows(e, raw, type, (int64_t)raw.size() / row_bytes(type, cols), cols, w);
}

// Dense down-projection.  K = ffn can exceed what one tensor-core GEMV holds in shared memory (x fragments of K > 16384:
// the 70B shape has ffn 28672); such a matrix is cut into column slabs that run side by side in one launch, exactly like
// the selected experts of a MoE layer, each on its own slice of x -- instead of falling back to the CUDA-core kernel,
// which streams these matrices at a quarter of the HBM rate (profiles/r02/r02_gemv_c4_shapes.log).
int upload_down(zb_engine* e, const GTensor* dn, DW& w, bool pairs_feed) {
    std::vector<uint8_t> raw;
    int type;
    int64_t cols;
    if (int rc = gather_rows({dn}, 0, -1, raw, type, cols)) return rc;
    const int64_t rows = (int64_t)raw.size() / row_bytes(type, cols);
    static const bool no_slabs = getenv("ZB_NO_KSLABS") && getenv("ZB_NO_KSLABS")[0] == '1';
    const int unit = type == kQ4_0 ? 128 : 256;
    const bool mma_type = type == kQ4_K || type == kQ5_K || type == kQ6_K || type == kQ4_0;
    if (!no_slabs && pairs_feed && e->use_mma && e->opts.batch <= 1 && e->tp_size == 1 && mma_type && rows % 16 == 0 && cols % unit == 0 &&
        zb_mma_check(type, (int)rows, (int)cols) != 0) {
        const int64_t nbk = cols / unit;
        int S = 0;
        for (int s = (int)((cols + 8191) / 8192); s <= 16; s++)
            if (nbk % s == 0 && zb_mma_check(type, (int)rows, (int)(cols / s)) == 0) { S = s; break; }
        if (S > 1) {
            const int64_t ks = cols / S, slab_bytes = rows * row_bytes(type, ks);
            std::vector<uint8_t> stacked((size_t)(slab_bytes * S));
            for (int sidx = 0; sidx < S; sidx++)
                if (zb_tp_shard_host(type, raw.data(), rows, cols, 0, rows, sidx * ks, (sidx + 1) * ks, stacked.data() + (size_t)sidx * slab_bytes))
                    return fail(ZB_EUNSUPPORTED, "down projection: cannot cut %lld columns into %d slabs", (long long)cols, S);
            if (int rc = upload_raw_rows(e, stacked, type, rows * S, ks, w, S, S)) return rc;
            if (w.e_mma_stride) {
                w.kslabs = S;
                e->max_kslabs = std::max(e->max_kslabs, S);
                return 0;
            }
            return fail(ZB_ESTATE, "down projection: slab stack was not given block-tiles");
        }
    }
    return upload_raw_rows(e, raw, type, rows, cols, w);
}

int upload_plain(zb_engine* e, const GTensor* t, DW& w) {  // bytes as they are in the file
    w.type = t->type; w.rows = t->rows(); w.cols = t->cols(); w.bytes = t->nbytes();
    uint8_t* d = nullptr;
    if (int rc = dalloc(e, &d, (size_t)w.bytes + 16)) return rc;
    CK(cudaMemcpy(d, t->data, (size_t)w.bytes, cudaMemcpyHostToDevice));
    w.d = d;
    return 0;
}

// Algorithmic bytes of one GEMV launch: weight blocks once + x + y.
double gemv_bytes(const DW& w, int nsel = 1) { return nsel * ((double)w.bytes + 4.0 * (double)w.cols + 4.0 * (double)w.rows); }

struct Sel {  // MoE expert indirection of one launch
    const int* idx = nullptr;
    int n = 0, a_stride = 0, y_stride = 0;
};

int gemv(zb_engine* e, const DW& w, const zb_prologue& p, float* y, bool pdl, const Sel& sel = Sel(), int xsite = -1);
void tp_site_producer(zb_engine* e, int site, zb_stream_weight& sw);

int gemv(zb_engine* e, const DW& w, const zb_prologue& p, float* y, bool pdl, const Sel& sel, int xsite) {
    zb_stream_weight sw{};
    sw.main = w.main; sw.aux = w.aux; sw.qtype = w.type; sw.rows = (int)w.rows; sw.cols = (int)w.cols;
    sw.epilogue = w.pairs ? 1 : 0;
    if (xsite >= 0) tp_site_producer(e, xsite, sw);
    zb_prologue pr = p;
    if (sel.idx) {
        sw.expert_sel = sel.idx; sw.n_sel = sel.n; sw.y_slot_stride = sel.y_stride;
        sw.expert_main_stride = w.e_main_stride; sw.expert_aux_stride = w.e_aux_stride;
        pr.a_slot_stride = sel.a_stride;
    }
    cudaStream_t s = e->stream;
    size_t i = 0;
    if (e->prof_on) {
        i = e->prof_rec.size() * 2;
        while (e->prof_ev.size() < i + 2) {
            cudaEvent_t ev;
            CK(cudaEventCreate(&ev));
            e->prof_ev.push_back(ev);
        }
        CK(cudaEventRecord(e->prof_ev[i], s));
    }
    int rc;
    if (w.mma && e->mma_scratch && (!sel.idx || w.e_mma_stride) && xsite < 0 && (p.mix_n == 0 || (!sel.idx && !p.swiglu)) && p.n_wait == 0) {
        zb_mma_weight mw{};
        mw.data = w.mma; mw.qtype = w.type; mw.rows = (int)w.rows; mw.cols = (int)w.cols; mw.epilogue = w.pairs ? 1 : 0;
        if (sel.idx) { mw.expert_sel = sel.idx; mw.n_sel = sel.n; mw.y_slot_stride = sel.y_stride; mw.expert_stride = w.e_mma_stride; }
        rc = zb_gemv_mma_f32(&mw, &pr, y, e->mma_scratch, (pdl && !e->prof_on) ? 1 : 0, (zb_stream_t)s);
    } else {
        rc = zb_gemv_stream_f32(&sw, &pr, y, (pdl && !e->prof_on) ? 1 : 0, (zb_stream_t)s);
    }
    if (rc) return fail(rc, "streamed gemv type %d [%lld x %lld]: %s", w.type, (long long)w.rows, (long long)w.cols, cudaGetErrorString((cudaError_t)rc));
    if (e->prof_on) {
        CK(cudaEventRecord(e->prof_ev[i + 1], s));
        e->prof_rec.push_back({w.type, gemv_bytes(w, sel.idx ? sel.n : 1)});
    }
    return 0;
}

// --------------------------------------------------------------------------
// Engine-private kernels
// --------------------------------------------------------------------------
// Token select + embedding row gather with bit-exact dequantisation (+ Gemma scale):
// prompt ids come from the feed buffer, afterwards the previous step's argmax
// (inference/arch_llama.go:246-342, arch_gemma.go:38).  Ids are clamped like
// launch_gather (gather.cu:20-23); the host API rejects out-of-range ids.
__global__ void embed_kernel(int type, const uint8_t* __restrict__ table, const int* __restrict__ feed, const int* __restrict__ feed_idx,
                             const int* __restrict__ feed_len, const int* __restrict__ last, float* __restrict__ out, int hidden, int vocab,
                             float scale) {
    int fi = *feed_idx;
    int tok = fi < *feed_len ? feed[fi] : *last;
    if (tok < 0) tok = 0;
    if (tok >= vocab) tok = vocab - 1;
    int64_t base = (int64_t)tok * hidden;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hidden; i += gridDim.x * blockDim.x) {
        float v = deq_raw(type, table, base + i);
        out[i] = scale > 0.0f ? v * scale : v;
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

__global__ void step_end_kernel(int* pos, int* feed_idx, const int* feed_len, const int* amax, int* last, int* out, int* n_out, int out_cap,
                                int with_head, int* step_count) {
    *pos += 1;
    if (step_count) *step_count += 1;  // epoch base of the fused tensor-parallel exchanges: never reset, identical on all ranks
    if (*feed_idx < *feed_len) *feed_idx += 1;
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
// lowest-index tie-break, weights renormalised to sum 1.
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

// Expert-sharded MoE: global expert id -> index in this rank's stack, or -1 (another rank computes that slot).
__global__ void moe_local_sel_kernel(const int* __restrict__ idx, int K, int experts_local, int rank, int* __restrict__ out) {
    int k = threadIdx.x;
    if (k < K) {
        int e = idx[k];
        out[k] = (e / experts_local == rank) ? e % experts_local : -1;
    }
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
    return upload_plain(e, t, w);
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
        if (e->top_k > e->n_experts) e->top_k = e->n_experts;
        // slots of one expert-indirect launch, the routing kernel's index buffers and zb_gemv_mma_f32 all assume a small top-k
        if (e->top_k < 1 || e->top_k > 16) return fail(ZB_EUNSUPPORTED, "expert_used_count %d outside 1..16", e->top_k);
    }
    if (is_gemma) e->embed_scale = (float)sqrt((double)e->hidden);
    if (is_gemma3) { e->post_norm = true; e->qk_norm = true; } else e->softcap = 0.0f;
    // arch_mistral.go:40, arch_mixtral.go:191, arch_starcoder2.go:42 hand cfg.SlidingWindow to every attention layer; the mask is
    // applied by the prompt pass only (grouped_query_attention.go:1074-1077)
    if (ar == "mistral" || ar == "mixtral" || ar == "starcoder2") e->prefill_window = (int)kvnum(g, ar, "attention.sliding_window", 0);
    if (is_moe && e->post_norm) return fail(ZB_EUNSUPPORTED, "MoE with post-norms is not supported");
    if (e->hidden <= 0 || e->layers <= 0 || e->n_q <= 0 || e->n_kv <= 0 || e->hd <= 0 || e->n_q % e->n_kv)
        return fail(ZB_EFORMAT, "invalid model dimensions (hidden %d layers %d heads %d/%d head_dim %d)", e->hidden, e->layers, e->n_q, e->n_kv, e->hd);
    // tensor parallel: this rank owns n_q/P query heads, n_kv/P KV heads (and their slice of the cache), ffn/P FFN columns
    const int P = e->tp_size, pr = e->tp_rank;
    Slicer sl;
    e->n_q_global = e->n_q;
    e->n_kv_global = e->n_kv;
    if (P > 1) {
        if (e->n_q % P || e->n_kv % P) return fail(ZB_EUNSUPPORTED, "tensor parallel %d does not divide heads %d/%d", P, e->n_q, e->n_kv);
        e->n_q /= P;
        e->n_kv /= P;
    }
    int rep = e->n_q / e->n_kv;
    if (!(e->hd == 32 || e->hd == 64 || e->hd == 128 || e->hd == 256)) return fail(ZB_EUNSUPPORTED, "head_dim %d unsupported (32, 64, 128, 256)", e->hd);
    if (!(rep == 1 || rep == 2 || rep == 3 || rep == 4 || rep == 8)) return fail(ZB_EUNSUPPORTED, "GQA ratio %d unsupported (1, 2, 3, 4, 8)", rep);
    e->max_seq = e->opts.max_seq > 0 ? e->opts.max_seq : (ctx < 4096 ? ctx : 4096);
    if (e->max_seq > ctx) e->max_seq = ctx;

    const GTensor *t_embed, *t_onorm;
    if (int rc = need(g, "token_embd.weight", &t_embed)) return rc;
    if (int rc = need(g, "output_norm.weight", &t_onorm)) return rc;
    e->vocab = (int)t_embed->rows();
    if (t_embed->cols() != e->hidden) return fail(ZB_EFORMAT, "token_embd row length %lld != hidden %d", (long long)t_embed->cols(), e->hidden);
    if (int rc = upload_plain(e, t_embed, e->embed_raw)) return rc;
    if (int rc = load_norm(e, "output_norm.weight", e->out_norm, true)) return rc;
    const GTensor* t_head = g.find("output.weight");
    if (!t_head) t_head = t_embed;  // tied head (arch_llama.go:54-58, arch_gemma.go:36)
    e->vocab_local = e->vocab;
    if (P > 1) {  // vocab rows split; logits shards meet in an all-gather
        if (e->vocab % P) return fail(ZB_EUNSUPPORTED, "tensor parallel %d does not divide the vocabulary %d", P, e->vocab);
        e->vocab_local = e->vocab / P;
        t_head = sl.rows(t_head, (int64_t)pr * e->vocab_local, (int64_t)(pr + 1) * e->vocab_local);
        if (!t_head) return fail(ZB_EFORMAT, "lm_head shard failed");
    }
    if (P == 1 && t_head == t_embed && (t_head->type == kQ4_K || t_head->type == kQ5_K)) {  // raw blocks already are the stream layout
        e->lm_head = e->embed_raw;
        e->lm_head.main = (uint8_t*)e->embed_raw.d;
        if (zb_stream_check(t_head->type, (int)t_head->rows(), (int)t_head->cols())) return fail(ZB_EUNSUPPORTED, "lm_head shape unsupported");
    } else if (int rc = upload(e, {t_head}, e->lm_head)) return rc;
    e->weight_bytes += e->lm_head.bytes;

    std::vector<float> cs, sn;
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
        if (P > 1) {
            const int64_t ql = (int64_t)e->n_q * e->hd, kl = (int64_t)e->n_kv * e->hd;
            q = sl.rows(q, pr * ql, (pr + 1) * ql);
            k = sl.rows(k, pr * kl, (pr + 1) * kl);
            v = sl.rows(v, pr * kl, (pr + 1) * kl);
            o = (o->cols() % P == 0) ? sl.cols(o, pr * (o->cols() / P), (pr + 1) * (o->cols() / P)) : nullptr;
            if (!q || !k || !v || !o) return fail(ZB_EUNSUPPORTED, "layer %d: attention weights cannot be sharded %d ways at block boundaries", i, P);
        }
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
            if (r->type != kF32) return fail(ZB_EUNSUPPORTED, "layer %d: router must be F32", i);
            if (int rc = upload_plain(e, r, L.router)) return rc;
            int E = e->n_experts;
            e->experts_local = E;
            if (P > 1) {  // experts sharded: rank p owns the contiguous block [p*E/P, (p+1)*E/P); the router stays replicated
                if (E % P) return fail(ZB_EUNSUPPORTED, "tensor parallel %d does not divide %d experts", P, E);
                const int El = E / P;
                const int64_t fr_g = ge->rows() / E, dr_g = de->rows() / E;
                ge = sl.rows(ge, (int64_t)pr * El * fr_g, (int64_t)(pr + 1) * El * fr_g);
                ue = sl.rows(ue, (int64_t)pr * El * fr_g, (int64_t)(pr + 1) * El * fr_g);
                de = sl.rows(de, (int64_t)pr * El * dr_g, (int64_t)(pr + 1) * El * dr_g);
                if (!ge || !ue || !de) return fail(ZB_EFORMAT, "layer %d: expert shard failed", i);
                E = El;
                e->experts_local = El;
            }
            if (ge->type != ue->type || ge->cols() != ue->cols() || ge->rows() != ue->rows() || ge->rows() % E || de->rows() % E)
                return fail(ZB_EUNSUPPORTED, "layer %d: expert tensors must share type and shape", i);
            // expert x occupies rows [x*fr, (x+1)*fr) of the stacked tensor (extractExpertSlice): build [gate_x ; up_x] per expert
            const int64_t fr = ge->rows() / E, rb = row_bytes(ge->type, ge->cols());
            std::vector<uint8_t> raw((size_t)(2 * fr * E * rb));
            for (int x = 0; x < E; x++)
                interleave_pairs(ge->data + (size_t)x * fr * rb, ue->data + (size_t)x * fr * rb, fr, rb, raw.data() + (size_t)(2 * x) * fr * rb);
            if (int rc = upload_raw_rows(e, raw, ge->type, 2 * fr * E, ge->cols(), L.e_gate_up, E)) return rc;
            L.e_gate_up.pairs = true;
            std::vector<uint8_t> rawd(de->data, de->data + de->nbytes());
            if (int rc = upload_raw_rows(e, rawd, de->type, de->rows(), de->cols(), L.e_down, E)) return rc;
            if (L.e_down.rows != e->hidden || L.e_down.cols != fr) return fail(ZB_EFORMAT, "layer %d: expert down shape mismatch", i);
            e->ffn = (int)fr;
            e->weight_bytes += L.router.bytes + (int64_t)e->top_k * (L.e_gate_up.bytes + L.e_down.bytes);
        } else {
            const GTensor *ga, *up, *dn;
            if (int rc = need(g, p + "ffn_gate.weight", &ga)) return rc;
            if (int rc = need(g, p + "ffn_up.weight", &up)) return rc;
            if (int rc = need(g, p + "ffn_down.weight", &dn)) return rc;
            if (P > 1) {
                if (ga->rows() % P) return fail(ZB_EUNSUPPORTED, "tensor parallel %d does not divide the FFN width %lld", P, (long long)ga->rows());
                const int64_t fl = ga->rows() / P;
                ga = sl.rows(ga, pr * fl, (pr + 1) * fl);
                up = sl.rows(up, pr * fl, (pr + 1) * fl);
                dn = sl.cols(dn, pr * fl, (pr + 1) * fl);
                if (!ga || !up || !dn) return fail(ZB_EUNSUPPORTED, "layer %d: FFN weights cannot be sharded %d ways at block boundaries", i, P);
            }
            if (ga->rows() != up->rows() || dn->cols() != ga->rows() || dn->rows() != e->hidden)
                return fail(ZB_EFORMAT, "layer %d: FFN weight shapes do not match", i);
            e->ffn = (int)ga->rows();
            if (ga->type == up->type && ga->cols() == up->cols()) {
                const int64_t rb = row_bytes(ga->type, ga->cols());
                std::vector<uint8_t> raw((size_t)(2 * ga->rows() * rb));
                interleave_pairs(ga->data, up->data, ga->rows(), rb, raw.data());
                DW w;
                if (int rc = upload_raw_rows(e, raw, ga->type, 2 * ga->rows(), ga->cols(), w)) return rc;
                w.pairs = true;
                L.gate_up.push_back(w);
            } else if (int rc = upload_group(e, {ga, up}, L.gate_up)) return rc;
            if (int rc = upload_down(e, dn, L.down, L.gate_up.size() == 1 && L.gate_up[0].pairs)) return rc;
            for (auto& w : L.gate_up) e->weight_bytes += w.bytes;
            e->weight_bytes += L.down.bytes * std::max(1, L.down.kslabs);
        }
        bool global = !(e->sw_pattern > 0 && ((i + 1) % e->sw_pattern != 0));  // arch_common.go:171-178
        L.cos_tbl = global ? e->tbl_gc : e->tbl_lc;
        L.sin_tbl = global ? e->tbl_gs : e->tbl_ls;
        size_t kvsz = (size_t)e->max_seq * e->n_kv * e->hd;
        if (e->opts.batch > 1) {  // paged pool shared by the sequences: [block][n_kv][page][hd] (generate/block_pool.go:36-67)
            int blocks_per_seq = (e->max_seq + 15) / 16;
            kvsz = (size_t)e->opts.batch * blocks_per_seq * e->n_kv * 16 * e->hd;
        }
        if (e->kv_f16) kvsz = (kvsz + 1) / 2;   // fp16 elements behind the float pointers
        if (int rc = dalloc(e, &L.kc, kvsz)) return rc;
        if (int rc = dalloc(e, &L.vc, kvsz)) return rc;
    }

    int qd = e->n_q * e->hd, kvd = e->n_kv * e->hd;
    int slots = is_moe ? e->top_k : 1;
    e->chunk = 32;
    {   // positions per attention CTA (multiple of the 16-position page): larger tiles trade split-merge latency for per-CTA work
        const char* ac = getenv("ZB_ATTN_CHUNK");
        int c = (ac && ac[0]) ? atoi(ac) : 0;
        if (c >= 16 && c % 16 == 0 && (size_t)c * e->hd * 8 <= 160 * 1024) e->chunk = c;
    }
    e->max_splits = (e->max_seq + e->chunk - 1) / e->chunk;
    {   // short contexts: the largest tile (multiple of 16 positions) whose K and V fit one CTA's shared memory next to the scratch
        const char* off = getenv("ZB_ATTN_SHORT");
        const int rep = e->n_q / e->n_kv;
        int c = (int)((200 * 1024 - (size_t)(rep + 16) * e->hd * 4 - 1024) / ((size_t)e->hd * 8)) / 16 * 16;
        if (c > e->max_seq) c = (e->max_seq + 15) / 16 * 16;
        const bool disabled = off && off[0] == '0';
        if (off && atoi(off) >= 64 && atoi(off) < c) c = atoi(off) / 16 * 16;   // ZB_ATTN_SHORT=n: smaller tile (tests cross the switch)
        e->chunk_short = (!disabled && c >= 64 && c > e->chunk && 16 * rep <= 2 * c && e->tp_size == 1 && e->opts.batch <= 1) ? c : 0;
    }
    if (int rc = dalloc(e, &e->hid, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->res, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->normed, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->qkv, qd + 2 * kvd)) return rc;
    if (int rc = dalloc(e, &e->attn, qd)) return rc;
    if (int rc = dalloc(e, &e->proj_o, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->proj, e->hidden)) return rc;
    if (int rc = dalloc(e, &e->gateup, 2 * (size_t)e->ffn * slots)) return rc;
    if (int rc = dalloc(e, &e->moe_y, (size_t)e->hidden * slots)) return rc;
    if (int rc = dalloc(e, &e->rlogits, 256)) return rc;
    if (int rc = dalloc(e, &e->rw, 256)) return rc;
    if (int rc = dalloc(e, &e->logits, e->vocab)) return rc;
    if (int rc = dalloc(e, &e->logits_local, e->vocab_local)) return rc;
    if (int rc = dalloc(e, &e->part_o, (size_t)e->n_q * e->max_splits * e->hd)) return rc;
    if (int rc = dalloc(e, &e->part_ml, 2 * (size_t)e->n_q * e->max_splits)) return rc;
    float* sc = nullptr;
    if (int rc = dalloc(e, &sc, 2 * (size_t)((e->vocab + 255) / 256) + 64)) return rc;
    e->amax_scratch = sc;
    e->feed_cap = e->max_seq;
    e->out_cap = e->max_seq;
    int* ints = nullptr;
    if (int rc = dalloc(e, &ints, 16 + 256 + 256 + (size_t)e->feed_cap + e->out_cap)) return rc;
    e->d_last = ints + 1; e->d_pos = ints + 2; e->d_feed_idx = ints + 4; e->d_feed_len = ints + 5;
    e->d_nout = ints + 6; e->d_amax = ints + 7;
    e->d_window_on = ints + 9;
    e->d_ridx = ints + 16;
    e->d_ridx_local = ints + 16 + 128;
    e->d_ticket = ints + 16 + 256;
    e->d_feed = ints + 16 + 512;
    e->d_out = e->d_feed + e->feed_cap;
    CK(cudaMallocHost(&e->h_pin, 64));
    return 0;
}

// --------------------------------------------------------------------------
// Program of the persistent whole-token kernel (zb_mega.cuh): the same op sequence as enqueue_step below, as a device-side
// table.  Eligible: dense model, one GPU, one sequence, every matrix in block-tiles.  Anything else keeps the CUDA graph.
// --------------------------------------------------------------------------
int mega_build(zb_engine* e) {
    e->mega = false;
    if (!e->want_mega || e->tp_size > 1 || e->opts.batch > 1 || e->n_experts > 0) return 0;
    const int rep = e->n_q / e->n_kv;
    if (!mega_attn_supported(e->hd, rep)) return 0;
    auto has_tiles = [&](const DW& w) { return w.mma != nullptr && !w.e_mma_stride; };
    bool ok = has_tiles(e->lm_head);
    for (auto& L : e->L) {
        for (auto& w : L.qkv) ok = ok && has_tiles(w);
        for (auto& w : L.gate_up) ok = ok && has_tiles(w);
        ok = ok && has_tiles(L.o) && has_tiles(L.down);
    }
    if (!ok) return 0;
    int ctas = 0;
    if (int rc = mega_max_ctas(e->opts.device, &ctas)) {
        cudaGetLastError();
        (void)rc;
        return 0;
    }
    std::vector<MegaOp> ops;
    std::vector<MegaStream> streams;
    int region = (int)((attn_item_floats(e->chunk, e->hd, rep, kMegaAttnWarps) * 4 + 127) & ~(size_t)127);
    int n_bar = 0;
    bool bad = false;
    // Ops whose epilogue needs finished sums (SwiGLU pairs, lm_head) split the matrix at row-tile boundaries; the others at
    // block-tile granularity, a straddling row tile being published as partial sums in planes (zb_mega.cuh MegaVec).
    auto geom_of = [&](const DW& w, bool whole, MGeom& g) { return make_mgeom(w.type, (int)w.rows, (int)w.cols, g, ctas, whole); };
    auto planes_of = [&](const DW& w) {
        MGeom g{};
        if (!geom_of(w, false, g)) { bad = true; return 1; }
        int mx = 1;
        for (int tt = 0; tt < g.n_tiles; tt++) mx = std::max(mx, ((tt + 1) * g.nb - 1) / g.per_cta - (tt * g.nb) / g.per_cta + 1);
        return mx;
    };
    int qkv_planes = 1, o_planes = 1, down_planes = 1;
    int64_t qkv_dim = 0, gu_dim = 0;
    for (auto& L : e->L) {
        int64_t q = 0, gu = 0;
        for (auto& w : L.qkv) { qkv_planes = std::max(qkv_planes, planes_of(w)); q += w.rows; }
        for (auto& w : L.gate_up) gu += w.rows;
        o_planes = std::max(o_planes, planes_of(L.o));
        down_planes = std::max(down_planes, planes_of(L.down));
        qkv_dim = std::max(qkv_dim, q);
        gu_dim = std::max(gu_dim, gu);
    }
    if (bad) return 0;
    auto round4 = [](int64_t n) { return (int)((n + 3) & ~(int64_t)3); };
    const int H4 = round4(e->hidden), QKV4 = round4(qkv_dim), GU4 = round4(gu_dim), AT4 = round4((int64_t)e->n_q * e->hd);
    uint2 *hid_ll = nullptr, *res_ll = nullptr, *qkv_ll = nullptr, *attn_ll = nullptr, *po_ll = nullptr, *gu_ll = nullptr, *proj_ll = nullptr;
    if (int rc = dalloc(e, &hid_ll, H4)) return rc;
    if (int rc = dalloc(e, &res_ll, H4)) return rc;
    if (int rc = dalloc(e, &qkv_ll, (size_t)QKV4 * qkv_planes)) return rc;
    if (int rc = dalloc(e, &attn_ll, AT4)) return rc;
    if (int rc = dalloc(e, &po_ll, (size_t)H4 * o_planes)) return rc;
    if (int rc = dalloc(e, &gu_ll, GU4)) return rc;
    if (int rc = dalloc(e, &proj_ll, (size_t)H4 * down_planes)) return rc;
    CK(cudaMemset(hid_ll, 0, (size_t)H4 * 8)); CK(cudaMemset(res_ll, 0, (size_t)H4 * 8)); CK(cudaMemset(qkv_ll, 0, (size_t)QKV4 * qkv_planes * 8));
    CK(cudaMemset(attn_ll, 0, (size_t)AT4 * 8)); CK(cudaMemset(po_ll, 0, (size_t)H4 * o_planes * 8)); CK(cudaMemset(gu_ll, 0, (size_t)GU4 * 8));
    CK(cudaMemset(proj_ll, 0, (size_t)H4 * down_planes * 8));

    struct Pro { MegaVec a, r; const float *w1, *w2; uint2* sum_out; float* sum_plain; int swiglu; };
    auto vec = [](const uint2* p, int planes, int stride, int tag) { return MegaVec{p, planes, stride, tag, 0}; };
    // y = flagged output (planes x stride) at element offset `off`; tag = index of the op whose epoch the vector carries
    auto push_gemv = [&](const DW& w, const Pro& p, uint2* y, int y_planes, int y_stride, int tag, bool barrier, bool head) {
        MGeom g{};
        const bool whole = head || w.pairs;
        if (!geom_of(w, whole, g)) { bad = true; return; }
        MegaOp op;
        memset(&op, 0, sizeof op);
        op.kind = kMegaGemv;
        op.barrier = barrier ? 1 : 0;
        MegaGemv& m = op.g;
        m.w = w.mma; m.y = y; m.y_plain = head ? e->logits : nullptr;
        m.a = p.a; m.r = p.r; m.w1 = p.w1; m.w2 = p.w2; m.sum_out = p.sum_out; m.sum_plain = p.sum_plain; m.eps = e->eps; m.swiglu = p.swiglu;
        m.tag_op = tag; m.y_planes = y_planes; m.y_stride = y_stride;
        m.type = w.type; m.M = (int)w.rows; m.K = (int)w.cols; m.pairs = w.pairs ? 1 : 0;
        m.nb = g.nb; m.n_tiles = g.n_tiles; m.total = g.total; m.per_cta = g.per_cta; m.per_warp = g.per_warp; m.slots = g.slots; m.max_local = g.max_local;
        m.xf_off = g.xf_off; m.xm_off = g.xm_off; m.xinv_off = g.xinv_off; m.part_off = g.part_off;
        m.stream = (int)streams.size();
        m.head = head ? 1 : 0;
        m.softcap = head ? e->softcap : 0.0f;
        if (g.ring_off > region) region = g.ring_off;
        streams.push_back(MegaStream{w.mma, g.total, g.per_cta, g.per_warp, bt_bytes(w.type), 0, 0});
        ops.push_back(op);
        if (barrier) n_bar++;
    };
    {
        MegaOp op;
        memset(&op, 0, sizeof op);
        op.kind = kMegaEmbed;
        op.barrier = 0;
        op.e = MegaEmbed{(const uint8_t*)e->embed_raw.d, e->d_feed, e->d_feed_idx, e->d_feed_len, e->d_last, hid_ll, e->embed_raw.type, e->hidden, e->vocab, e->embed_scale};
        ops.push_back(op);
    }
    // the residual stream alternates between hid_ll and res_ll; `cur` is where it is (or will be after the pending prologue)
    Pro pend{};
    pend.a = vec(hid_ll, 1, H4, 0);   // written by the embed op (index 0)
    uint2* cur = hid_ll;
    int cur_tag = 0;
    for (int li = 0; li < e->layers; li++) {
        Layer& L = e->L[li];
        Pro pq = pend;
        pq.w2 = (const float*)L.attn_norm.d;
        const int qkv_tag = (int)ops.size();   // the group's first op: every op of the group tags its rows with it
        int64_t off = 0;
        for (size_t i = 0; i < L.qkv.size(); i++) {
            Pro p1 = pq;
            if (i) { p1.sum_out = nullptr; p1.sum_plain = nullptr; }
            // sum_out (CTA 0, first op of the group) carries the first op's tag = qkv_tag as well
            push_gemv(L.qkv[i], p1, qkv_ll + off, qkv_planes, QKV4, qkv_tag, false, false);
            off += L.qkv[i].rows;
        }
        if (pend.sum_out) { cur = pend.sum_out; cur_tag = qkv_tag; }
        const int attn_op = (int)ops.size();
        {
            MegaOp op;
            memset(&op, 0, sizeof op);
            op.kind = kMegaAttn;
            op.barrier = 0;
            op.a = AttnArgs{nullptr, e->qk_norm ? (const float*)L.q_norm.d : nullptr, e->qk_norm ? (const float*)L.k_norm.d : nullptr, L.cos_tbl, L.sin_tbl,
                            e->d_pos, L.kc, L.vc, nullptr, e->part_o, e->part_ml, e->d_ticket, e->eps, (float)(1.0 / sqrt((double)e->hd)), e->hd, e->n_q,
                            e->n_kv, e->max_seq, e->chunk, e->max_splits, nullptr, 0, 16, 0, 0, kMegaAttnWarps, qkv_ll, qkv_planes, QKV4, qkv_tag, attn_ll,
                            e->prefill_window, e->d_window_on};
            ops.push_back(op);
        }
        Pro po{};
        po.a = vec(attn_ll, 1, AT4, attn_op);
        const int o_op = (int)ops.size();
        push_gemv(L.o, po, po_ll, o_planes, H4, o_op, false, false);
        uint2* other = cur == hid_ll ? res_ll : hid_ll;
        Pro pf{};
        pf.a = vec(po_ll, o_planes, H4, o_op);
        pf.w1 = e->post_norm ? (const float*)L.post_attn_norm.d : nullptr;
        pf.r = vec(cur, 1, H4, cur_tag);
        pf.sum_out = other;
        pf.w2 = (const float*)L.ffn_norm.d;
        const int gu_tag = (int)ops.size();
        off = 0;
        for (size_t i = 0; i < L.gate_up.size(); i++) {
            Pro p1 = pf;
            if (i) { p1.sum_out = nullptr; p1.sum_plain = nullptr; }
            push_gemv(L.gate_up[i], p1, gu_ll + off, 1, GU4, gu_tag, false, false);
            off += L.gate_up[i].pairs ? L.gate_up[i].rows / 2 : L.gate_up[i].rows;
        }
        cur = other;
        cur_tag = gu_tag;
        Pro pd{};
        pd.a = vec(gu_ll, 1, GU4, gu_tag);
        pd.swiglu = L.gate_up[0].pairs ? 0 : 1;
        const int down_op = (int)ops.size();
        push_gemv(L.down, pd, proj_ll, down_planes, H4, down_op, false, false);
        pend = Pro{};
        pend.a = vec(proj_ll, down_planes, H4, down_op);
        pend.w1 = e->post_norm ? (const float*)L.post_ffw_norm.d : nullptr;
        pend.r = vec(cur, 1, H4, cur_tag);
        pend.sum_out = cur == hid_ll ? res_ll : hid_ll;
    }
    const int n_streams_nohead = (int)streams.size();
    float* mega_final_hid = e->hid;   // plain copy of the post-stack residual stream for the host (zb_engine_hidden)
    {
        Pro ph = pend;
        ph.w2 = (const float*)e->out_norm.d;
        ph.sum_out = nullptr;            // nothing after the head reads the stream
        ph.sum_plain = mega_final_hid;
        push_gemv(e->lm_head, ph, nullptr, 1, 0, (int)ops.size(), true, true);
    }
    if (bad) return 0;
    // the CTA's ring: slots of the largest block-tile of the model, as many as fit after the scratch region (+ 2 mbarriers each)
    int slot_bytes = 0;
    for (auto& st : streams) slot_bytes = std::max(slot_bytes, (st.bt + 127) & ~127);
    region = (region + 127) & ~127;
    const int nslots = slot_bytes > 0 ? (kMegaSmem - region - 64) / (slot_bytes + 16) : 0;
    if (nslots < 2 * kMW) return 0;   // activation fragments of a very wide matrix leave no room to stream: keep the graph path
    int* mints = nullptr;
    if (int rc = dalloc(e, &mints, 64)) return rc;
    CK(cudaMemset(mints, 0, 64 * sizeof(int)));
    e->d_mega_bar = reinterpret_cast<unsigned int*>(mints);
    {
        MegaOp op;
        memset(&op, 0, sizeof op);
        op.kind = kMegaFinal;
        op.barrier = 0;
        op.f = MegaFinal{e->d_pos, e->d_feed_idx, e->d_amax, e->d_last, e->d_out, e->d_nout, mints + 1, reinterpret_cast<unsigned int*>(mints + 8), e->d_feed_len, e->out_cap};
        ops.push_back(op);
    }
    MegaOp* d_ops = nullptr;
    MegaStream* d_streams = nullptr;
    float* cand_v = nullptr;
    int* cand_i = nullptr;
    if (int rc = dalloc(e, &d_ops, ops.size())) return rc;
    if (int rc = dalloc(e, &d_streams, streams.size())) return rc;
    if (int rc = dalloc(e, &cand_v, ZB_SMS)) return rc;
    if (int rc = dalloc(e, &cand_i, ZB_SMS)) return rc;
    CK(cudaMemcpy(d_ops, ops.data(), ops.size() * sizeof(MegaOp), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_streams, streams.data(), streams.size() * sizeof(MegaStream), cudaMemcpyHostToDevice));
    MegaCtl& c = e->mctl;
    c.ops = d_ops; c.streams = d_streams;
    c.n_ops = (int)ops.size(); c.n_streams = (int)streams.size(); c.n_streams_nohead = n_streams_nohead;
    c.n_barriers = n_bar; c.region_bytes = region; c.nslots = nslots; c.slot_bytes = slot_bytes;
    c.bar_counter = e->d_mega_bar; c.step = mints + 1;
    c.epoch_step = reinterpret_cast<unsigned int*>(mints + 8);   // outside the 8 bytes zb_engine_reset clears: epochs never repeat
    c.cand_v = cand_v; c.cand_i = cand_i;
    c.trace = nullptr;
    {   // ZB_MEGA_TRACE=1: per-op, per-CTA phase stamps of the last launch (zb_engine_mega_trace)
        const char* tr = getenv("ZB_MEGA_TRACE");
        if (tr && tr[0] && strcmp(tr, "0")) {
            long long* t = nullptr;
            if (int rc = dalloc(e, &t, ops.size() * (size_t)ctas * kMegaTraceSlots)) return rc;
            c.trace = t;
        }
    }
    e->mega_ctas = ctas;
    e->final_hid = mega_final_hid;
    e->mega = true;
    return 0;
}

// --------------------------------------------------------------------------
// One decode step, enqueued on the engine stream (capturable).
//
// Per layer (dense): 5 launches, each a programmatic dependent of the one before
// it so its weight ring fills while its predecessor computes:
//   QKV  GEMV  x = RMSNorm(residual)            [residual add of the previous layer fused in]
//   attention stage (QK-norm, RoPE, KV append, split-KV decode, merge)
//   O    GEMV
//   gate|up GEMV  x = RMSNorm(o (+post-attn norm) + residual)   [fusedAddRMSNormNode]
//   down GEMV  x = SwiGLU(gate, up)
// against the reference's ~12 nodes per layer.
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

// Wait-only consumer of a fused exchange (steps without lm_head still have to retire the last site's flags).
__global__ void tp_wait_kernel(const uint2* slots, int n, int stride, const int* epoch_base, int site, int sites_per_step) {
    if ((int)threadIdx.x < n) {  // element 0 of every peer's slot carrying this epoch proves that peer has retired the previous site
        const unsigned int want = (unsigned int)(epoch_base[0] * sites_per_step + site + 1);
        unsigned int v, ep, spins = 0;
        do {
            asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v), "=r"(ep) : "l"(slots + (size_t)threadIdx.x * stride) : "memory");
            if (++spins > (1u << 24)) __trap();
        } while (ep != want);
    }
}

// One exchange buffer per rank, mapped into every peer with cudaIpc (handles travel through the NCCL communicator that
// already exists).  Row-parallel GEMVs then push their partial sums straight into the peers' slots over NVLink and the
// consuming prologue adds the P slots in rank order: GEMV + all-reduce in one kernel pair, no separate collective launch.
int tp_setup_fused(zb_engine* e) {
    const int P = e->tp_size;
    const char* off = getenv("ZB_TP_NCCL_ONLY");
    if (P > 8 || e->n_experts > 0 || (off && off[0] && strcmp(off, "0"))) return 0;  // MoE keeps the NCCL all-reduce
    // Two users of the mapped exchange buffers.  (a) Fused: the row-parallel GEMV's epilogue pushes its partial sums into the
    // peers' slots and every consumer CTA adds the P slots in its prologue -- no extra launch, but P * hidden re-reads per CTA,
    // so it only pays while P * hidden is small (measured on 8xB200, 70B shape, 4 layers: P=8 0.706 vs 0.589 ms/step with NCCL;
    // hidden 512 at TP=2: 0.138 vs 0.185).  (b) One-shot push all-reduce (tp_allreduce_push_kernel): the GEMVs stay on the
    // tensor-core kernel, one small launch pushes the rank's vector to every peer as flagged pairs and sums the P slots once
    // into a plain vector -- the default for everything (a) does not cover.  ZB_TP_FUSED=1 / ZB_TP_PUSH=0 force the choice.
    const char* force = getenv("ZB_TP_FUSED");
    const bool want_fused = (force && force[0]) ? strcmp(force, "0") != 0 : (long long)P * e->hidden <= 8192;   // ZB_TP_FUSED=0: never
    const char* push = getenv("ZB_TP_PUSH");
    const bool want_push = !want_fused && !(push && push[0] && !strcmp(push, "0"));
    if (!want_fused && !want_push) return 0;
    e->xn = (e->hidden + 63) & ~63;
    e->xchg_slots_off = 1024;
    size_t bytes = e->xchg_slots_off + (size_t)2 * P * e->xn * 8;  // (value, epoch) pairs
    if (int rc = dalloc(e, &e->xchg_local, bytes)) return rc;
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, e->xchg_local));
    uint8_t *d_mine = nullptr, *d_all = nullptr;
    if (int rc = dalloc(e, &d_mine, sizeof mine)) return rc;
    if (int rc = dalloc(e, &d_all, sizeof mine * P)) return rc;
    CK(cudaMemcpy(d_mine, &mine, sizeof mine, cudaMemcpyHostToDevice));
    NCCLK(g_nccl.AllGather(d_mine, d_all, sizeof mine, 0 /*ncclInt8*/, e->nccl_comm, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    std::vector<cudaIpcMemHandle_t> all(P);
    CK(cudaMemcpy(all.data(), d_all, sizeof mine * P, cudaMemcpyDeviceToHost));
    for (int p = 0; p < P; p++) {
        if (p == e->tp_rank) { e->xchg_peer[p] = e->xchg_local; continue; }
        void* ptr = nullptr;
        cudaError_t ce = cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess) {  // no peer mapping (e.g. no NVLink / different process namespace): stay on NCCL
            cudaGetLastError();
            for (int q = 0; q < p; q++)
                if (q != e->tp_rank && e->xchg_peer[q]) { cudaIpcCloseMemHandle(e->xchg_peer[q]); e->xchg_peer[q] = nullptr; }
            return 0;
        }
        e->xchg_peer[p] = (uint8_t*)ptr;
    }
    std::vector<float> ones(8, 1.0f);
    if (int rc = dalloc(e, &e->d_ones, 8)) return rc;
    CK(cudaMemcpy(e->d_ones, ones.data(), 32, cudaMemcpyHostToDevice));
    int* ints = nullptr;
    if (int rc = dalloc(e, &ints, 8)) return rc;
    e->d_step = ints;
    e->d_xticket = ints + 1;
    // every rank must have mapped everybody before anyone pushes: a tiny all-reduce is the barrier
    NCCLK(g_nccl.AllReduce(e->d_ones, e->d_ones, 1, 7, 2 /*ncclMax*/, e->nccl_comm, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->tp_fused = want_fused;
    e->tp_push = want_push;
    return 0;
}

// producer / consumer descriptors of exchange site `site` (parity = site & 1)
void tp_site_producer(zb_engine* e, int site, zb_stream_weight& sw) {
    const int P = e->tp_size, par = site & 1;
    sw.n_peers = P; sw.site = site; sw.sites_per_step = 2 * e->layers;
    for (int p = 0; p < P; p++) {
        sw.peer_out[p] = reinterpret_cast<float*>(reinterpret_cast<uint2*>(e->xchg_peer[p] + e->xchg_slots_off) + (size_t)(par * P + e->tp_rank) * e->xn);
        sw.peer_flag[p] = reinterpret_cast<unsigned int*>(e->xchg_peer[p]) + par * P + e->tp_rank;
    }
    sw.epoch_base = e->d_step;
    sw.ticket = e->d_xticket;
}
void tp_site_consumer(zb_engine* e, int site, zb_prologue& p) {
    const int P = e->tp_size, par = site & 1;
    p.a = reinterpret_cast<const float*>(reinterpret_cast<const uint2*>(e->xchg_local + e->xchg_slots_off) + (size_t)par * P * e->xn);
    p.mix_w = e->d_ones; p.mix_n = P; p.mix_stride = e->xn;
    p.wait_flags = reinterpret_cast<const unsigned int*>(e->xchg_local) + par * P;
    p.wait_epoch_base = e->d_step;
    p.n_wait = P; p.wait_site = site; p.wait_sites_per_step = 2 * e->layers;
}

// One-shot all-reduce of a hidden-size vector over NVLink peer memory, latency-bound by design (32 KB per rank): every rank
// writes its partial vector as flagged (value, epoch) pairs -- 16-byte stores, two elements each -- into slot [rank] of every
// peer's exchange buffer, then polls its own P slots and adds them in rank order (identical sums on every rank).  Data and
// "ready" travel in the same store (the LL idea), so there is no barrier, no second hop and no reduction tree: one NVLink
// write latency plus one local L2 poll.  Replaces ncclAllReduce after o_proj / down_proj
// (inference/parallel/tensor_parallel.go:151-163, distributed/nccl.go:90-98) when the peers are mapped.
struct PushPeers { uint2* slot[8]; };   // slot [rank] of the site's parity in every peer's buffer (own buffer included)

// `in` and `out` may be the same buffer (the engine reduces in place): every thread reads its two inputs before it writes them.
__global__ void __launch_bounds__(256) tp_allreduce_push_kernel(const float* in, float* out, int n, PushPeers peers,
                                                                const uint2* __restrict__ local, int P, int xn, const int* __restrict__ epoch_base,
                                                                int site, int sites_per_step) {
    const unsigned int epoch = (unsigned int)(*epoch_base) * (unsigned int)sites_per_step + (unsigned int)site + 1u;
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    const float2 v = *reinterpret_cast<const float2*>(in + i);
    for (int p = 0; p < P; p++)
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(peers.slot[p] + i), "r"(__float_as_uint(v.x)), "r"(epoch),
                     "r"(__float_as_uint(v.y)), "r"(epoch)
                     : "memory");
    float s0 = 0.0f, s1 = 0.0f;
    for (int q = 0; q < P; q++) {
        const uint2* src = local + (size_t)q * xn + i;
        unsigned int a, fa, b, fb, spins = 0;
        do {
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(src) : "memory");
            if (++spins > (1u << 26)) __trap();   // a lost peer traps instead of hanging the GPU
        } while (fa != epoch || fb != epoch);
        s0 = q == 0 ? __uint_as_float(a) : s0 + __uint_as_float(a);
        s1 = q == 0 ? __uint_as_float(b) : s1 + __uint_as_float(b);
    }
    *reinterpret_cast<float2*>(out + i) = make_float2(s0, s1);
}

__global__ void step_bump_kernel(int* step) { *step += 1; }

// Tensor-parallel exchange on the engine stream (graph-capturable): the sum of the row-parallel partials
// (inference/parallel/tensor_parallel.go:151-163 AllReduceSum) after o_proj and down_proj.
int tp_allreduce(zb_engine* e, float* buf, size_t n, Counter& cnt) {
    if (e->tp_size <= 1) return 0;
    if (e->tp_push && (int)n <= e->xn && !(n & 1)) {
        const int P = e->tp_size, site = e->ar_site++, par = site & 1;
        PushPeers pp{};
        for (int p = 0; p < P; p++)
            pp.slot[p] = reinterpret_cast<uint2*>(e->xchg_peer[p] + e->xchg_slots_off) + (size_t)(par * P + e->tp_rank) * e->xn;
        const uint2* local = reinterpret_cast<const uint2*>(e->xchg_local + e->xchg_slots_off) + (size_t)par * P * e->xn;
        const int threads = (int)(n / 2);
        KLAUNCH(tp_allreduce_push_kernel<<<(threads + 255) / 256, 256, 0, e->stream>>>(buf, buf, (int)n, pp, local, P, e->xn, e->d_step, site,
                                                                                      e->ar_sites_per_step));
        return 0;
    }
    NCCLK(g_nccl.AllReduce(buf, buf, n, 7 /*ncclFloat32*/, 0 /*ncclSum*/, e->nccl_comm, e->stream));
    cnt.n++;
    return 0;
}

int enqueue_step(zb_engine* e, bool with_head, Counter& cnt) {
    cudaStream_t s = e->stream;
    e->ar_site = 0;
    e->ar_sites_per_step = 2 * e->layers;
    const int H = e->hidden, hd = e->hd, nq = e->n_q, nkv = e->n_kv;
    const bool pdl = e->use_pdl && !e->prof_on;
    KLAUNCH(embed_kernel<<<(H + 255) / 256, 256, 0, s>>>(e->embed_raw.type, (const uint8_t*)e->embed_raw.d, e->d_feed, e->d_feed_idx, e->d_feed_len,
                                                         e->d_last, e->hid, H, e->vocab, e->embed_scale));
    // `pend` describes the not-yet-materialised residual stream: v = [rmsnorm(a, w1) | mix(a)] + r, stored to sum_out by the next prologue
    zb_prologue pend{};
    pend.a = e->hid;
    pend.eps = e->eps;
    const float* cur = e->hid;  // where the residual stream is (or will be after the next prologue)
    for (int li = 0; li < e->layers; li++) {
        Layer& L = e->L[li];
        // ---- attention block
        zb_prologue pq = pend;
        pq.w2 = (const float*)L.attn_norm.d;
        int64_t off = 0;
        bool first = true;
        for (auto& w : L.qkv) {
            zb_prologue p1 = pq;
            if (!first) p1.sum_out = nullptr;
            if (int rc = gemv(e, w, p1, e->qkv + off, pdl)) return rc;
            cnt.n++;
            off += w.rows;
            first = false;
        }
        if (pend.sum_out) cur = pend.sum_out;
        zb_attn_args aa{};
        aa.qkv = e->qkv;
        aa.q_norm = e->qk_norm ? (const float*)L.q_norm.d : nullptr;
        aa.k_norm = e->qk_norm ? (const float*)L.k_norm.d : nullptr;
        aa.cos_tbl = L.cos_tbl; aa.sin_tbl = L.sin_tbl; aa.pos = e->d_pos;
        aa.k_cache = L.kc; aa.v_cache = L.vc; aa.out = e->attn; aa.part_o = e->part_o; aa.part_ml = e->part_ml; aa.ticket = e->d_ticket;
        aa.eps = e->eps; aa.head_dim = hd; aa.n_q = nq; aa.n_kv = nkv; aa.max_seq = e->max_seq; aa.chunk = e->chunk; aa.max_splits = e->max_splits;
        aa.window = e->prefill_window; aa.window_on = e->d_window_on;
        aa.kv_f16 = e->kv_f16 ? 1 : 0;
        int aflags = pdl ? 1 : 0;
        aa.warps = 8 * (nq / nkv) <= 2 * e->chunk ? 8 : 4;   // measured: 8 warps per 32-position tile 786 vs 769 tok/s on C2
        if (e->attn_short) { aa.chunk = e->chunk_short; aa.max_splits = 1; aa.warps = 16; aflags |= ZB_ATTN_SINGLE_TILE; }
        LAUNCH(zb_decode_attn_f32(&aa, aflags, (zb_stream_t)s));
        zb_prologue po{};
        po.a = e->attn;
        po.eps = e->eps;
        const bool fused = e->tp_fused && !L.router.d;
        if (int rc = gemv(e, L.o, po, e->proj_o, pdl, Sel(), fused ? 2 * li : -1)) return rc;
        cnt.n++;
        if (!fused)
            if (int rc = tp_allreduce(e, e->proj_o, H, cnt)) return rc;
        // ---- FFN block: residual + pre-FFN norm fused into the gate|up prologue (fusedAddRMSNormNode)
        float* other = cur == e->hid ? e->res : e->hid;
        zb_prologue pf{};
        pf.a = e->proj_o;
        if (fused) tp_site_consumer(e, 2 * li, pf);   // a = sum over ranks of the o_proj partials, straight from the exchange slots
        pf.w1 = e->post_norm ? (const float*)L.post_attn_norm.d : nullptr;  // Gemma 3 (arch_common.go:404-416)
        pf.r = cur;
        pf.sum_out = other;
        pf.w2 = (const float*)L.ffn_norm.d;
        pf.eps = e->eps;
        pend = zb_prologue{};
        pend.eps = e->eps;
        if (L.router.d) {
            // MoE (layers/core/moe.go:74-160,405-553): the router needs the normed vector in memory
            LAUNCH(fused_add_rmsnorm_f32(e->proj_o, cur, (const float*)L.ffn_norm.d, e->normed, other, f2u(e->eps), 1, H, s));
            LAUNCH(launch_sgemv_m1(e->rlogits, (const float*)L.router.d, e->normed, e->n_experts, H, s));
            KLAUNCH(moe_route_kernel<<<1, 32, 0, s>>>(e->rlogits, e->n_experts, e->top_k, e->d_ridx, e->rw));
            zb_prologue pg{};
            pg.a = e->normed;
            pg.eps = e->eps;
            const int* sel = e->d_ridx;
            if (e->tp_size > 1) {
                KLAUNCH(moe_local_sel_kernel<<<1, 32, 0, s>>>(e->d_ridx, e->top_k, e->experts_local, e->tp_rank, e->d_ridx_local));
                sel = e->d_ridx_local;
            }
            Sel sg{sel, e->top_k, 0, e->ffn};   // SwiGLU applied in the epilogue: slot k writes act[k][ffn]
            if (int rc = gemv(e, L.e_gate_up, pg, e->gateup, pdl, sg)) return rc;
            cnt.n++;
            zb_prologue pd{};
            pd.a = e->gateup;
            pd.eps = e->eps;
            Sel sd{sel, e->top_k, e->ffn, H};
            if (int rc = gemv(e, L.e_down, pd, e->moe_y, pdl, sd)) return rc;
            cnt.n++;
            if (int rc = tp_allreduce(e, e->moe_y, (size_t)e->top_k * H, cnt)) return rc;
            pend.a = e->moe_y;
            pend.mix_w = e->rw;
            pend.mix_n = e->top_k;
            pend.mix_stride = H;
        } else {
            off = 0;
            first = true;
            for (auto& w : L.gate_up) {
                zb_prologue p1 = pf;
                if (!first) p1.sum_out = nullptr;
                if (int rc = gemv(e, w, p1, e->gateup + off, pdl)) return rc;
                cnt.n++;
                off += w.rows;
                first = false;
            }
            zb_prologue pd{};
            pd.a = e->gateup;
            pd.swiglu = L.gate_up[0].pairs ? 0 : 1;  // pairs: the gate|up epilogue already wrote silu(gate)*up
            pd.eps = e->eps;
            if (L.down.kslabs > 1) {   // column slabs side by side; the next prologue adds the partial outputs in slab order
                Sel sd{e->d_iota, L.down.kslabs, (int)L.down.cols, H};
                if (int rc = gemv(e, L.down, pd, e->slab_y, pdl, sd)) return rc;
                cnt.n++;
                pend.a = e->slab_y;
                pend.mix_w = e->d_fones;
                pend.mix_n = L.down.kslabs;
                pend.mix_stride = H;
            } else {
            if (int rc = gemv(e, L.down, pd, e->proj, pdl, Sel(), fused ? 2 * li + 1 : -1)) return rc;
            cnt.n++;
            if (!fused)
                if (int rc = tp_allreduce(e, e->proj, H, cnt)) return rc;
            pend.a = e->proj;
            }
            if (fused) tp_site_consumer(e, 2 * li + 1, pend);
            pend.w1 = e->post_norm ? (const float*)L.post_ffw_norm.d : nullptr;  // fusedNormAddNode (Gemma 3) / residual add
        }
        cur = other;
        pend.r = cur;
        pend.sum_out = cur == e->hid ? e->res : e->hid;
    }
    if (with_head) {
        e->final_hid = pend.sum_out ? pend.sum_out : (pend.n_wait ? e->hid : pend.a);
        zb_prologue ph = pend;
        ph.w2 = (const float*)e->out_norm.d;
        if (int rc = gemv(e, e->lm_head, ph, e->tp_size > 1 ? e->logits_local : e->logits, pdl)) return rc;
        cnt.n++;
        if (e->tp_size > 1) {  // vocab shards -> full logits on every rank (rank order = row order)
            NCCLK(g_nccl.AllGather(e->logits_local, e->logits, (size_t)e->vocab_local, 7, e->nccl_comm, e->stream));
            cnt.n++;
        }
        if (e->softcap > 0.0f)
            KLAUNCH(softcap_kernel<<<(e->vocab + 255) / 256, 256, 0, s>>>(e->logits, e->vocab, e->softcap, (float)(1.0 / (double)e->softcap)));
        LAUNCH(launch_argmax(e->logits, e->d_amax, e->amax_scratch, e->vocab, s));
        cnt.n++;  // two stages
    }
    if (e->tp_fused && !with_head && pend.n_wait > 0)  // nobody consumed the last down_proj exchange in a step without lm_head
        KLAUNCH(tp_wait_kernel<<<1, 32, 0, s>>>(reinterpret_cast<const uint2*>(pend.a), pend.n_wait, pend.mix_stride, pend.wait_epoch_base, pend.wait_site,
                                                pend.wait_sites_per_step));
    KLAUNCH(step_end_kernel<<<1, 1, 0, s>>>(e->d_pos, e->d_feed_idx, e->d_feed_len, e->d_amax, e->d_last, e->d_out, e->d_nout, e->out_cap,
                                            with_head ? 1 : 0, e->d_step));
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
    if (e->mega) {   // one cooperative launch per token
        int rc = mega_launch(e->mctl, e->mega_ctas, with_head ? 1 : 0, e->stream);
        if (rc) return fail(rc, "decode_mega_kernel launch: %s", cudaGetErrorString((cudaError_t)rc));
        if (with_head) e->launches_full = 1;
        e->host_pos++;
        return 0;
    }
    const bool shortc = e->chunk_short > 0 && e->host_pos + 1 <= e->chunk_short;
    cudaGraphExec_t gx = with_head ? (shortc && e->graph_full_s ? e->graph_full_s : e->graph_full)
                                   : (shortc && e->graph_nohead_s ? e->graph_nohead_s : e->graph_nohead);
    if (gx) {
        CK(cudaGraphLaunch(gx, e->stream));
    } else {
        Counter cnt;
        e->attn_short = shortc;
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
    if (int rc = mega_build(e)) return rc;
    if (int rc = zb_engine_reset(e)) return rc;
    if (int rc = run_step(e, true)) return rc;
    CK(cudaStreamSynchronize(e->stream));
    if (e->mega) {   // warm the variant without lm_head as well; nothing to capture
        if (int rc = run_step(e, false)) return rc;
        CK(cudaStreamSynchronize(e->stream));
        return zb_engine_reset(e);
    }
    if (e->opts.use_graph) {
        e->attn_short = false;
        int rc = capture(e, true, &e->graph_full);
        if (rc && e->use_pdl) {  // a driver that cannot capture programmatic edges: capture plain launches instead
            cudaGetLastError();
            e->use_pdl = false;
            rc = capture(e, true, &e->graph_full);
        }
        if (rc) return rc;
        if (int rc2 = capture(e, false, &e->graph_nohead)) return rc2;
        if (e->chunk_short > 0) {
            e->attn_short = true;
            int rs = capture(e, true, &e->graph_full_s);
            if (!rs) rs = capture(e, false, &e->graph_nohead_s);
            e->attn_short = false;
            if (rs) return rs;
        }
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


// ==========================================================================
// Batched decode (BASELINE config 3: B = 32 over a paged KV cache).
// The reference has no batched forward (BatchGenerate is sequential, the graph input is [1, seqLen], the
// paged cache is host memory that re-gathers O(T) per step); this is the device-resident equivalent behind the same
// step API: B sequences advance in lock-step, KV lives in 16-position blocks taken from a shared pool through a
// per-sequence block table (generate/paged_kv.go:74-185, block_pool.go:36-67), every matmul is the tcgen05 GEMM.
// ==========================================================================
__global__ void embed_batch_kernel(int type, const uint8_t* __restrict__ table, const int* __restrict__ toks, float* __restrict__ out, int hidden,
                                   int vocab, float scale) {
    int tok = toks[blockIdx.y];
    if (tok < 0) tok = 0;
    if (tok >= vocab) tok = vocab - 1;
    int64_t base = (int64_t)tok * hidden;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hidden; i += gridDim.x * blockDim.x) {
        float v = deq_raw(type, table, base + i);
        out[(size_t)blockIdx.y * hidden + i] = scale > 0.0f ? v * scale : v;
    }
}

// one CTA per sequence: optional Gemma softcap, then argmax with the lowest index winning ties (argmax.cu:40-48)
__global__ void argmax_rows_kernel(float* __restrict__ logits, int n, float cap, int* __restrict__ result) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float* row = logits + (size_t)blockIdx.x * n;
    float bv = -3.402823466e38f;
    int bi = 0x7fffffff;
    const float inv_cap = cap > 0.0f ? (float)(1.0 / (double)cap) : 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = row[i];
        if (cap > 0.0f) {
            float x = v * inv_cap, t;
            if (x > 4.5f) t = 1.0f;
            else if (x < -4.5f) t = -1.0f;
            else { float x2 = x * x; t = x * (27.0f + x2) / (27.0f + 9.0f * x2); }
            v = cap * t;
            row[i] = v;
        }
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sv[w] = bv; si[w] = bi; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        bv = l < nw ? sv[l] : -3.402823466e38f;
        bi = l < nw ? si[l] : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (l == 0) result[blockIdx.x] = bi;
    }
}

__global__ void step_end_batch_kernel(int* pos, int* tok, const int* amax, int* out, int* n_out, int out_cap, int B) {
    int b = threadIdx.x;
    if (b >= B) return;
    pos[b] += 1;
    int t = amax[b];
    tok[b] = t;
    int n = *n_out;
    if (n < out_cap) out[(size_t)n * B + b] = t;
    __syncthreads();
    if (b == 0) *n_out = n + 1;
}

bool is_kquant(int t) { return t == kQ4_K || t == kQ5_K || t == kQ6_K; }

int batch_alloc(zb_engine* e) {
    const int B = e->opts.batch;
    e->B = B;
    e->Bpad = (B + 15) / 16 * 16;
    e->max_blocks = (e->max_seq + e->page - 1) / e->page;
    e->pool_blocks = B * e->max_blocks;
    if (e->n_experts > 0) return fail(ZB_EUNSUPPORTED, "batched decode of MoE models is not supported yet");
    if (B > 256) return fail(ZB_EUNSUPPORTED, "batch %d > 256", B);
    auto chk = [&](const DW& w) { return is_kquant(w.type) && w.cols % 256 == 0; };
    bool ok = chk(e->lm_head);
    for (auto& L : e->L) {
        for (auto& w : L.qkv) ok = ok && chk(w);
        for (auto& w : L.gate_up) ok = ok && chk(w);
        ok = ok && chk(L.o) && chk(L.down);
    }
    if (!ok) return fail(ZB_EUNSUPPORTED, "batched decode needs K-quant (Q4_K/Q5_K/Q6_K) matrices with K %% 256 == 0");
    const int H = e->hidden, qd = e->n_q * e->hd, kvd = e->n_kv * e->hd;
    int kmax = std::max(std::max(H, qd), e->ffn);
    if (int rc = dalloc(e, &e->b_hid, (size_t)B * H)) return rc;
    if (int rc = dalloc(e, &e->b_res, (size_t)B * H)) return rc;
    if (int rc = dalloc(e, &e->b_qkv, (size_t)B * (qd + 2 * kvd))) return rc;
    if (int rc = dalloc(e, &e->b_attn, (size_t)B * qd)) return rc;
    if (int rc = dalloc(e, &e->b_proj_o, (size_t)B * H)) return rc;
    if (int rc = dalloc(e, &e->b_proj, (size_t)B * H)) return rc;
    if (int rc = dalloc(e, &e->b_gateup, (size_t)B * 2 * e->ffn)) return rc;
    if (int rc = dalloc(e, &e->b_logits, (size_t)B * e->vocab)) return rc;
    uint16_t* xb = nullptr;
    if (int rc = dalloc(e, &xb, (size_t)e->Bpad * kmax)) return rc;
    e->b_xhi = xb;
    if (int rc = dalloc(e, &xb, (size_t)e->Bpad * kmax)) return rc;
    e->b_xlo = xb;
    float* po = nullptr;  // attention scratch per sequence (the single-sequence buffers are too small)
    if (int rc = dalloc(e, &po, (size_t)B * e->n_q * e->max_splits * e->hd)) return rc;
    e->part_o = po;
    if (int rc = dalloc(e, &po, (size_t)B * 2 * e->n_q * e->max_splits)) return rc;
    e->part_ml = po;
    int* ints = nullptr;
    if (int rc = dalloc(e, &ints, (size_t)B * (e->max_blocks + 4 + e->n_kv) + 16 + (size_t)e->out_cap * B)) return rc;
    e->d_btab = ints;
    e->d_bpos = ints + (size_t)B * e->max_blocks;
    e->d_btok = e->d_bpos + B;
    e->d_bamax = e->d_btok + B;
    e->d_bnout = e->d_bamax + B;
    e->d_ticket = e->d_bnout + 16;
    e->d_bout = e->d_ticket + (size_t)B * e->n_kv;
    e->h_btab.assign((size_t)B * e->max_blocks, 0);
    e->h_bpos.assign(B, 0);
    CK(cudaMallocHost(&e->h_bpin, (size_t)B * 2 * sizeof(int)));
    return 0;
}

int bgemm(zb_engine* e, const DW& w, int K, float* y, int ldy, Counter& cnt) {
    zb_stream_weight sw{};
    sw.main = w.main; sw.aux = w.aux; sw.qtype = w.type; sw.rows = (int)w.rows; sw.cols = (int)w.cols;
    int rc = zb_gemm_tc_f32(&sw, e->b_xhi, e->B <= 64 ? e->b_xlo : nullptr, e->B, (e->B + 15) / 16 * 16, y, ldy, (zb_stream_t)e->stream);
    cnt.n++;
    if (rc) return fail(rc, "tcgen05 gemm type %d [%lld x %lld]: %s", w.type, (long long)w.rows, (long long)w.cols, cudaGetErrorString((cudaError_t)rc));
    return 0;
}

int bprep(zb_engine* e, zb_prep_args a, int K, int qtype, Counter& cnt) {
    a.K = K; a.qtype = qtype; a.eps = e->eps;
    a.xhi = e->b_xhi; a.xlo = e->B <= 64 ? e->b_xlo : nullptr; a.ldx = (e->B + 15) / 16 * 16;
    int rc = zb_gemm_tc_prep_rows(&a, e->B, (zb_stream_t)e->stream);
    cnt.n++;
    if (rc) return fail(rc, "batched prologue: %s", cudaGetErrorString((cudaError_t)rc));
    return 0;
}

int enqueue_batch_step(zb_engine* e, Counter& cnt) {
    cudaStream_t s = e->stream;
    const int B = e->B, H = e->hidden, hd = e->hd, nq = e->n_q, nkv = e->n_kv, qd = nq * hd, kvd = nkv * hd, F = e->ffn;
    KLAUNCH(embed_batch_kernel<<<dim3((H + 255) / 256, B), 256, 0, s>>>(e->embed_raw.type, (const uint8_t*)e->embed_raw.d, e->d_btok, e->b_hid, H,
                                                                       e->vocab, e->embed_scale));
    zb_prep_args pend{};
    pend.a = e->b_hid; pend.lda = H;
    const float* cur = e->b_hid;
    for (int li = 0; li < e->layers; li++) {
        Layer& L = e->L[li];
        zb_prep_args pq = pend;
        pq.w2 = (const float*)L.attn_norm.d;
        if (int rc = bprep(e, pq, H, L.qkv[0].type, cnt)) return rc;
        if (pend.sum_out) cur = pend.sum_out;
        int64_t off = 0;
        int prev_type = L.qkv[0].type;
        for (auto& w : L.qkv) {
            if (w.type != prev_type && ((w.type == kQ6_K) != (prev_type == kQ6_K))) {  // Q6_K consumes x in a different k-slot order
                zb_prep_args p2{};
                p2.a = cur; p2.lda = H; p2.w2 = (const float*)L.attn_norm.d;
                if (int rc = bprep(e, p2, H, w.type, cnt)) return rc;
            }
            prev_type = w.type;
            if (int rc = bgemm(e, w, H, e->b_qkv + off, qd + 2 * kvd, cnt)) return rc;
            off += w.rows;
        }
        zb_attn_args aa{};
        aa.qkv = e->b_qkv;
        aa.q_norm = e->qk_norm ? (const float*)L.q_norm.d : nullptr;
        aa.k_norm = e->qk_norm ? (const float*)L.k_norm.d : nullptr;
        aa.cos_tbl = L.cos_tbl; aa.sin_tbl = L.sin_tbl; aa.pos = e->d_bpos;
        aa.k_cache = L.kc; aa.v_cache = L.vc; aa.out = e->b_attn; aa.part_o = e->part_o; aa.part_ml = e->part_ml; aa.ticket = e->d_ticket;
        aa.eps = e->eps; aa.head_dim = hd; aa.n_q = nq; aa.n_kv = nkv; aa.max_seq = e->max_seq; aa.chunk = e->chunk; aa.max_splits = e->max_splits;
        aa.batch = B; aa.qkv_stride = qd + 2 * kvd; aa.out_stride = qd; aa.block_table = e->d_btab; aa.max_blocks = e->max_blocks; aa.page = e->page;
        aa.kv_f16 = e->kv_f16 ? 1 : 0;
        LAUNCH(zb_decode_attn_f32(&aa, 0, (zb_stream_t)s));
        zb_prep_args po{};
        po.a = e->b_attn; po.lda = qd;
        if (int rc = bprep(e, po, qd, L.o.type, cnt)) return rc;
        if (int rc = bgemm(e, L.o, qd, e->b_proj_o, H, cnt)) return rc;
        float* other = cur == e->b_hid ? e->b_res : e->b_hid;
        zb_prep_args pf{};
        pf.a = e->b_proj_o; pf.lda = H;
        pf.w1 = e->post_norm ? (const float*)L.post_attn_norm.d : nullptr;
        pf.r = cur; pf.ldr = H;
        pf.sum_out = other; pf.ldsum = H;
        pf.w2 = (const float*)L.ffn_norm.d;
        if (int rc = bprep(e, pf, H, L.gate_up[0].type, cnt)) return rc;
        off = 0;
        for (auto& w : L.gate_up) {
            if (w.type != L.gate_up[0].type) return fail(ZB_EUNSUPPORTED, "batched decode: gate and up must share a type");
            if (int rc = bgemm(e, w, H, e->b_gateup + off, 2 * F, cnt)) return rc;
            off += w.rows;
        }
        zb_prep_args pd{};
        pd.a = e->b_gateup; pd.lda = 2 * F;
        pd.mode = L.gate_up[0].pairs ? 1 : 2;
        if (int rc = bprep(e, pd, F, L.down.type, cnt)) return rc;
        if (int rc = bgemm(e, L.down, F, e->b_proj, H, cnt)) return rc;
        cur = other;
        pend = zb_prep_args{};
        pend.a = e->b_proj; pend.lda = H;
        pend.w1 = e->post_norm ? (const float*)L.post_ffw_norm.d : nullptr;
        pend.r = cur; pend.ldr = H;
        pend.sum_out = cur == e->b_hid ? e->b_res : e->b_hid; pend.ldsum = H;
    }
    zb_prep_args ph = pend;
    ph.w2 = (const float*)e->out_norm.d;
    if (int rc = bprep(e, ph, H, e->lm_head.type, cnt)) return rc;
    if (int rc = bgemm(e, e->lm_head, H, e->b_logits, e->vocab, cnt)) return rc;
    KLAUNCH(argmax_rows_kernel<<<B, 256, 0, s>>>(e->b_logits, e->vocab, e->softcap, e->d_bamax));
    KLAUNCH(step_end_batch_kernel<<<1, 256, 0, s>>>(e->d_bpos, e->d_btok, e->d_bamax, e->d_bout, e->d_bnout, e->out_cap, B));
    return 0;
}

// Give every sequence a block for the position it is about to write (BlockPool.Alloc, block_pool.go:36-67).
int batch_ensure_blocks(zb_engine* e) {
    bool dirty = false;
    for (int b = 0; b < e->B; b++) {
        int pos = e->h_bpos[b];
        if (pos >= e->max_seq) return fail(ZB_ESTATE, "KV cache full (%d positions)", e->max_seq);
        if (pos % e->page == 0) {
            if (e->free_blocks.empty()) return fail(ZB_ESTATE, "KV block pool exhausted");
            e->h_btab[(size_t)b * e->max_blocks + pos / e->page] = e->free_blocks.back();
            e->free_blocks.pop_back();
            dirty = true;
        }
    }
    if (dirty) CK(cudaMemcpyAsync(e->d_btab, e->h_btab.data(), e->h_btab.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    return 0;
}

int batch_run_step(zb_engine* e) {
    if (int rc = batch_ensure_blocks(e)) return rc;
    if (e->graph_batch) {
        CK(cudaGraphLaunch(e->graph_batch, e->stream));
    } else {
        Counter cnt;
        if (int rc = enqueue_batch_step(e, cnt)) return rc;
        e->launches_batch = cnt.n;
    }
    for (int b = 0; b < e->B; b++) e->h_bpos[b]++;
    return 0;
}

int batch_warm_and_capture(zb_engine* e) {
    if (int rc = zb_engine_batch_reset(e)) return rc;
    if (int rc = batch_run_step(e)) return rc;
    CK(cudaStreamSynchronize(e->stream));
    if (e->opts.use_graph) {
        cudaGraph_t graph = nullptr;
        Counter cnt;
        CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_batch_step(e, cnt);
        cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail((int)ce, "cudaStreamEndCapture (batch): %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&e->graph_batch, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return fail((int)ce, "cudaGraphInstantiate (batch): %s", cudaGetErrorString(ce));
        e->launches_batch = cnt.n;
    }
    return zb_engine_batch_reset(e);
}


// ==========================================================================
// Chunked prefill through the tcgen05 GEMMs (BASELINE config 3: 4k-token prompt).
// The reference prefills with one graph forward over [1, seqLen] (generate/session.go:150-160) whose matmuls hit
// dequant_q4k_f32 + cuBLAS SGEMM and flash_attention_forward_f32; here every matmul of a 256-token chunk is one tcgen05
// dequant-GEMM and attention is causal over the decode path's own KV cache, so decode continues on the same cache.
// ==========================================================================
struct ChunkBufs {
    int T = 0;
    float *hid = nullptr, *res = nullptr, *qkv = nullptr, *qrot = nullptr, *attn = nullptr, *proj_o = nullptr, *proj = nullptr, *gateup = nullptr,
          *xnorm = nullptr;
    void *xhi = nullptr, *xlo = nullptr;
};

int chunk_alloc(zb_engine* e, ChunkBufs& c, int T) {
    const int H = e->hidden, qd = e->n_q * e->hd, kvd = e->n_kv * e->hd;
    const int kmax = std::max(std::max(H, qd), e->ffn), tp = (T + 15) / 16 * 16;
    c.T = T;
    if (int rc = dalloc(e, &c.hid, (size_t)T * H)) return rc;
    if (int rc = dalloc(e, &c.res, (size_t)T * H)) return rc;
    if (int rc = dalloc(e, &c.qkv, (size_t)T * (qd + 2 * kvd))) return rc;
    if (int rc = dalloc(e, &c.qrot, (size_t)T * qd)) return rc;
    if (int rc = dalloc(e, &c.attn, (size_t)T * qd)) return rc;
    if (int rc = dalloc(e, &c.proj_o, (size_t)T * H)) return rc;
    if (int rc = dalloc(e, &c.proj, (size_t)T * H)) return rc;
    if (int rc = dalloc(e, &c.gateup, (size_t)T * 2 * e->ffn)) return rc;
    if (int rc = dalloc(e, &c.xnorm, (size_t)T * H)) return rc;
    uint16_t* xb = nullptr;
    if (int rc = dalloc(e, &xb, (size_t)tp * kmax)) return rc;
    c.xhi = xb;
    if (int rc = dalloc(e, &xb, (size_t)tp * kmax)) return rc;
    c.xlo = xb;
    return 0;
}

int cgemm(zb_engine* e, ChunkBufs& c, const DW& w, int K, int T, float* y, int ldy) {
    zb_stream_weight sw{};
    sw.main = w.main; sw.aux = w.aux; sw.qtype = w.type; sw.rows = (int)w.rows; sw.cols = (int)w.cols;
    int rc = zb_gemm_tc_f32(&sw, c.xhi, T <= 64 ? c.xlo : nullptr, T, (c.T + 15) / 16 * 16, y, ldy, (zb_stream_t)e->stream);
    if (rc) return fail(rc, "tcgen05 gemm type %d [%lld x %lld] x %d tokens: %s", w.type, (long long)w.rows, (long long)w.cols, T, cudaGetErrorString((cudaError_t)rc));
    return 0;
}

int cprep(zb_engine* e, ChunkBufs& c, zb_prep_args a, int K, int qtype, int T) {
    a.K = K; a.qtype = qtype; a.eps = e->eps;
    a.xhi = c.xhi; a.xlo = T <= 64 ? c.xlo : nullptr; a.ldx = (c.T + 15) / 16 * 16;
    int rc = zb_gemm_tc_prep_rows(&a, T, (zb_stream_t)e->stream);
    if (rc) return fail(rc, "chunk prologue: %s", cudaGetErrorString((cudaError_t)rc));
    return 0;
}

// One chunk of T prompt tokens (ids already in e->d_feed[off .. off+T)) at positions p0 .. p0+T-1.
// Leaves the post-stack residual description in `pend` (rows = tokens of the chunk).
int enqueue_chunk(zb_engine* e, ChunkBufs& c, int off, int T, int p0, zb_prep_args& pend) {
    cudaStream_t s = e->stream;
    const int H = e->hidden, hd = e->hd, nq = e->n_q, nkv = e->n_kv, qd = nq * hd, kvd = nkv * hd, F = e->ffn;
    embed_batch_kernel<<<dim3((H + 255) / 256, T), 256, 0, s>>>(e->embed_raw.type, (const uint8_t*)e->embed_raw.d, e->d_feed + off, c.hid, H, e->vocab,
                                                                 e->embed_scale);
    CK(cudaGetLastError());
    pend = zb_prep_args{};
    pend.a = c.hid; pend.lda = H;
    const float* cur = c.hid;
    for (int li = 0; li < e->layers; li++) {
        Layer& L = e->L[li];
        zb_prep_args pq = pend;
        pq.w2 = (const float*)L.attn_norm.d;
        if (int rc = cprep(e, c, pq, H, L.qkv[0].type, T)) return rc;
        if (pend.sum_out) cur = pend.sum_out;
        int64_t o2 = 0;
        int prev_type = L.qkv[0].type;
        for (auto& w : L.qkv) {
            if (w.type != prev_type && ((w.type == kQ6_K) != (prev_type == kQ6_K))) {
                zb_prep_args p2{};
                p2.a = cur; p2.lda = H; p2.w2 = (const float*)L.attn_norm.d;
                if (int rc = cprep(e, c, p2, H, w.type, T)) return rc;
            }
            prev_type = w.type;
            if (int rc = cgemm(e, c, w, H, T, c.qkv + o2, qd + 2 * kvd)) return rc;
            o2 += w.rows;
        }
        int rc = zb_prefill_attn_f32(c.qkv, qd + 2 * kvd, e->qk_norm ? (const float*)L.q_norm.d : nullptr, e->qk_norm ? (const float*)L.k_norm.d : nullptr,
                                     L.cos_tbl, L.sin_tbl, p0, T, c.qrot, L.kc, L.vc, c.attn, e->eps, hd, nq, nkv, e->max_seq, e->prefill_window, 0,
                                     (zb_stream_t)s);
        if (rc) return fail(rc, "prefill attention: %s", cudaGetErrorString((cudaError_t)rc));
        zb_prep_args po{};
        po.a = c.attn; po.lda = qd;
        if (int rc2 = cprep(e, c, po, qd, L.o.type, T)) return rc2;
        if (int rc2 = cgemm(e, c, L.o, qd, T, c.proj_o, H)) return rc2;
        float* other = cur == c.hid ? c.res : c.hid;
        zb_prep_args pf{};
        pf.a = c.proj_o; pf.lda = H;
        pf.w1 = e->post_norm ? (const float*)L.post_attn_norm.d : nullptr;
        pf.r = cur; pf.ldr = H;
        pf.sum_out = other; pf.ldsum = H;
        pf.w2 = (const float*)L.ffn_norm.d;
        if (int rc2 = cprep(e, c, pf, H, L.gate_up[0].type, T)) return rc2;
        o2 = 0;
        for (auto& w : L.gate_up) {
            if (w.type != L.gate_up[0].type) return fail(ZB_EUNSUPPORTED, "chunked prefill: gate and up must share a type");
            if (int rc2 = cgemm(e, c, w, H, T, c.gateup + o2, 2 * F)) return rc2;
            o2 += w.rows;
        }
        zb_prep_args pd{};
        pd.a = c.gateup; pd.lda = 2 * F;
        pd.mode = L.gate_up[0].pairs ? 1 : 2;
        if (int rc2 = cprep(e, c, pd, F, L.down.type, T)) return rc2;
        if (int rc2 = cgemm(e, c, L.down, F, T, c.proj, H)) return rc2;
        cur = other;
        pend = zb_prep_args{};
        pend.a = c.proj; pend.lda = H;
        pend.w1 = e->post_norm ? (const float*)L.post_ffw_norm.d : nullptr;
        pend.r = cur; pend.ldr = H;
        pend.sum_out = cur == c.hid ? c.res : c.hid; pend.ldsum = H;
    }
    return 0;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
ZB_API const char* zb_last_error(void) { return g_err.c_str(); }

ZB_API int zb_tp_unique_id(void* out128) {
    if (!out128) return fail(ZB_EINVAL, "zb_tp_unique_id: null argument");
    if (int rc = nccl_load()) return rc;
    NCCLK(g_nccl.GetUniqueId(static_cast<NcclId*>(out128)));
    return 0;
}

// Pure host: rows [r0, r1) x columns [c0, c1) of a raw GGUF matrix (column bounds at block boundaries).
ZB_API int zb_tp_shard_host(int qtype, const void* raw, int64_t rows, int64_t cols, int64_t r0, int64_t r1, int64_t c0, int64_t c1, void* out) {
    return shard_bytes(qtype, static_cast<const uint8_t*>(raw), rows, cols, r0, r1, c0, c1, static_cast<uint8_t*>(out)) ? ZB_EINVAL : 0;
}

static int engine_create_impl(const char* gguf_path, const zb_engine_opts* opts, const void* nccl_id, zb_engine** out);

ZB_API int zb_engine_create(const char* gguf_path, const zb_engine_opts* opts, zb_engine** out) {
    return engine_create_impl(gguf_path, opts, nullptr, out);
}

// One process per GPU: every rank passes the same 128-byte id (zb_tp_unique_id on rank 0, broadcast by the host).
ZB_API int zb_engine_create_tp(const char* gguf_path, const zb_engine_opts* opts, const void* nccl_id128, zb_engine** out) {
    if (!opts || opts->tp_size < 2 || !nccl_id128) return fail(ZB_EINVAL, "zb_engine_create_tp: needs opts.tp_size >= 2 and the NCCL unique id");
    return engine_create_impl(gguf_path, opts, nccl_id128, out);
}

static int engine_create_impl(const char* gguf_path, const zb_engine_opts* opts, const void* nccl_id, zb_engine** out) {
    if (!gguf_path || !out) return fail(ZB_EINVAL, "zb_engine_create: null argument");
    *out = nullptr;
    zb_engine* e = new zb_engine();
    if (opts) e->opts = *opts;
    else { e->opts.use_graph = 1; }
    if (e->opts.tp_size <= 0) e->opts.tp_size = 1;
    e->tp_size = e->opts.tp_size;
    e->tp_rank = e->opts.tp_rank;
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
        if (e->tp_size > 1) {
            if (!nccl_id) { rc = fail(ZB_EUNSUPPORTED, "tensor parallel engine: use zb_engine_create_tp"); break; }
            if (e->tp_rank < 0 || e->tp_rank >= e->tp_size) { rc = fail(ZB_EINVAL, "tp_rank %d out of range", e->tp_rank); break; }
            if (e->opts.batch > 1) { rc = fail(ZB_EUNSUPPORTED, "tensor parallel + batched decode is not supported yet"); break; }
            if ((rc = nccl_load())) break;
            NcclId id;
            memcpy(&id, nccl_id, sizeof id);
            int nr = g_nccl.CommInitRank(&e->nccl_comm, e->tp_size, id, e->tp_rank);
            if (nr) { rc = fail(ZB_EIO, "ncclCommInitRank failed: %d", nr); break; }
        }
        if (e->opts.tp_size != 1 && !e->nccl_comm) { rc = fail(ZB_EUNSUPPORTED, "tensor parallel engine: use zb_engine_create_tp"); break; }
        { const char* np = getenv("ZB_NO_PDL"); if (np && np[0] && strcmp(np, "0")) e->use_pdl = false; }
        { const char* tc = getenv("ZB_GEMV_TC"); if (tc && tc[0] && !strcmp(tc, "0")) e->use_mma = false; }
        {   // persistent whole-token kernel: opt-in (opts.flags & ZB_ENGINE_MEGA, or ZB_MEGA=1); the CUDA-graph step is the default
            const char* mg = getenv("ZB_MEGA");
            const bool asked = (e->opts.flags & ZB_ENGINE_MEGA) || (mg && mg[0] && strcmp(mg, "0"));
            e->want_mega = asked && e->use_mma && !(e->opts.flags & ZB_ENGINE_NO_MEGA) && e->tp_size == 1 && e->opts.batch <= 1;
        }
        {   // fp16 KV cache: opts.flags & ZB_ENGINE_KV_F16, or ZB_KV_F16=1
            const char* kf = getenv("ZB_KV_F16");
            e->kv_f16 = (e->opts.flags & ZB_ENGINE_KV_F16) || (kf && kf[0] && strcmp(kf, "0"));
            if (e->kv_f16) e->want_mega = false;   // the persistent kernel keeps the f32 cache
        }
        rc = load_model(e, gguf_path);
        if (rc) break;
        if (e->mma_scratch_bytes) {
            uint8_t* sp = nullptr;
            if ((rc = dalloc(e, &sp, (size_t)e->mma_scratch_bytes))) break;
            e->mma_scratch = sp;
        }
        if (e->max_kslabs > 1) {
            int iota[16];
            float ones[16];
            for (int k = 0; k < 16; k++) { iota[k] = k; ones[k] = 1.0f; }
            if ((rc = dalloc(e, &e->slab_y, (size_t)e->max_kslabs * e->hidden))) break;
            if ((rc = dalloc(e, &e->d_iota, 16))) break;
            if ((rc = dalloc(e, &e->d_fones, 16))) break;
            if (cudaMemcpy(e->d_iota, iota, sizeof iota, cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMemcpy(e->d_fones, ones, sizeof ones, cudaMemcpyHostToDevice) != cudaSuccess) { rc = fail(ZB_EIO, "slab tables: copy failed"); break; }
        }
        if (e->tp_size > 1 && (rc = tp_setup_fused(e))) break;
        if (e->opts.batch > 1) {
            rc = batch_alloc(e);
            if (!rc) rc = batch_warm_and_capture(e);
        } else {
            rc = warm_and_capture(e);
        }
    } while (0);
    if (rc) {
        delete e;
        return rc;
    }
    *out = e;
    return 0;
}

ZB_API void zb_engine_destroy(zb_engine* e) {
    if (e) delete static_cast<ChunkBufs*>(e->pchunk);
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
    o->kv_bytes_per_pos = 2LL * e->layers * e->n_kv * e->hd * (e->kv_f16 ? 2 : 4);
    o->launches_per_step = e->B > 1 ? e->launches_batch : e->launches_full;
    snprintf(o->arch, sizeof o->arch, "%s", e->arch.c_str());
    return 0;
}

ZB_API int zb_engine_reset(zb_engine* e) {
    if (!e) return fail(ZB_EINVAL, "null engine");
    CK(cudaSetDevice(e->opts.device));
    // ints: [1] last token, [2] position, [4] feed index, [5] feed length, [6] tokens out, [7] argmax; tickets re-armed
    CK(cudaMemsetAsync(e->d_last - 1, 0, (16 + 512) * sizeof(int), e->stream));
    if (e->d_mega_bar) CK(cudaMemsetAsync(e->d_mega_bar, 0, 8, e->stream));   // barrier arrivals and launch count restart together
    CK(cudaStreamSynchronize(e->stream));
    e->host_pos = 0;
    return 0;
}

ZB_API int zb_engine_prefill(zb_engine* e, const int32_t* tokens, int n, int32_t* first_token) {
    if (e && e->B > 1) return fail(ZB_ESTATE, "this engine was created with batch %d: use the zb_engine_batch_* entry points (the single-sequence API would read the paged KV pool as a contiguous cache)", e->B);
    if (!e || !tokens || n <= 0) return fail(ZB_EINVAL, "zb_engine_prefill: bad arguments");
    CK(cudaSetDevice(e->opts.device));
    if (e->host_pos + n > e->max_seq) return fail(ZB_ESTATE, "prompt does not fit the KV cache (%d + %d > %d)", e->host_pos, n, e->max_seq);
    if (int rc = set_feed(e, tokens, n)) return rc;
    const bool windowed = e->prefill_window > 0 && e->d_window_on;
    if (windowed) CK(cudaMemsetAsync(e->d_window_on, 1, 1, e->stream));   // the prompt pass masks beyond the sliding window
    for (int i = 0; i < n; i++)
        if (int rc = run_step(e, i == n - 1)) { if (windowed) cudaMemsetAsync(e->d_window_on, 0, 4, e->stream); return rc; }
    if (windowed) CK(cudaMemsetAsync(e->d_window_on, 0, 4, e->stream));
    CK(cudaMemcpyAsync(e->h_pin + 1, e->d_last, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (first_token) *first_token = e->h_pin[1];
    return 0;
}

ZB_API int zb_engine_prefill_chunked(zb_engine* e, const int32_t* tokens, int n, int32_t* first_token, float* ms) {
    if (!e || !tokens || n <= 0) return fail(ZB_EINVAL, "zb_engine_prefill_chunked: bad arguments");
    if (e->B > 1 || e->tp_size > 1 || e->n_experts > 0) return fail(ZB_EUNSUPPORTED, "chunked prefill: single-sequence dense engine without tensor parallelism only");
    if (e->kv_f16) return fail(ZB_EUNSUPPORTED, "chunked prefill writes an f32 KV cache; this engine stores fp16 (use zb_engine_prefill)");
    if (e->max_kslabs > 1) return fail(ZB_EUNSUPPORTED, "chunked prefill: the down projection of this model is stored as column slabs (ffn > 16384); set ZB_NO_KSLABS=1");
    CK(cudaSetDevice(e->opts.device));
    if (e->host_pos + n > e->max_seq) return fail(ZB_ESTATE, "prompt does not fit the KV cache (%d + %d > %d)", e->host_pos, n, e->max_seq);
    auto chk = [&](const DW& w) { return is_kquant(w.type) && w.cols % 256 == 0; };
    bool ok = true;
    for (auto& L : e->L) {
        for (auto& w : L.qkv) ok = ok && chk(w);
        for (auto& w : L.gate_up) ok = ok && chk(w);
        ok = ok && chk(L.o) && chk(L.down);
    }
    if (!ok) return fail(ZB_EUNSUPPORTED, "chunked prefill needs K-quant (Q4_K/Q5_K/Q6_K) matrices with K %% 256 == 0");
    const int TC = 256;
    if (!e->pchunk) {
        ChunkBufs* c = new ChunkBufs();
        if (int rc = chunk_alloc(e, *c, TC)) { delete c; return rc; }
        e->pchunk = c;
    }
    ChunkBufs& c = *static_cast<ChunkBufs*>(e->pchunk);
    if (int rc = set_feed(e, tokens, n)) return rc;
    CK(cudaEventRecord(e->ev0, e->stream));
    zb_prep_args pend{};
    int last_T = 0;
    for (int off = 0; off < n; off += TC) {
        const int T = std::min(TC, n - off);
        if (int rc = enqueue_chunk(e, c, off, T, e->host_pos + off, pend)) return rc;
        last_T = T;
    }
    // logits of the last prompt token: final residual + output norm for the chunk's rows (f32), then the batch-1 lm_head GEMV
    zb_prep_args ph = pend;
    ph.w2 = (const float*)e->out_norm.d;
    ph.x_f32 = c.xnorm; ph.ldxf = e->hidden;
    ph.K = e->hidden; ph.qtype = e->lm_head.type; ph.eps = e->eps;
    ph.xhi = nullptr; ph.xlo = nullptr;
    {
        int rc = zb_gemm_tc_prep_rows(&ph, last_T, (zb_stream_t)e->stream);
        if (rc) return fail(rc, "chunk final norm: %s", cudaGetErrorString((cudaError_t)rc));
    }
    zb_prologue pl{};
    pl.a = c.xnorm + (size_t)(last_T - 1) * e->hidden;
    pl.eps = e->eps;
    if (int rc = gemv(e, e->lm_head, pl, e->logits, false)) return rc;
    if (e->softcap > 0.0f) softcap_kernel<<<(e->vocab + 255) / 256, 256, 0, e->stream>>>(e->logits, e->vocab, e->softcap, (float)(1.0 / (double)e->softcap));
    CK((cudaError_t)launch_argmax(e->logits, e->d_amax, e->amax_scratch, e->vocab, e->stream));
    // bookkeeping: position += n, feed consumed, last token = argmax, one output token recorded
    int hdr[2] = {n, n};
    CK(cudaMemcpyAsync(e->d_feed_idx, hdr, 8, cudaMemcpyHostToDevice, e->stream));
    int newpos = e->host_pos + n;
    CK(cudaMemcpyAsync(e->d_pos, &newpos, 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->d_last, e->d_amax, 4, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(e->d_out, e->d_amax, 4, cudaMemcpyDeviceToDevice, e->stream));
    int one = 1;
    CK(cudaMemcpyAsync(e->d_nout, &one, 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaMemcpyAsync(e->h_pin + 1, e->d_last, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->host_pos = newpos;
    e->final_hid = nullptr;
    if (first_token) *first_token = e->h_pin[1];
    if (ms) CK(cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return 0;
}

ZB_API int zb_engine_decode_step(zb_engine* e, int32_t token, int32_t* next_token) {
    if (e && e->B > 1) return fail(ZB_ESTATE, "this engine was created with batch %d: use the zb_engine_batch_* entry points (the single-sequence API would read the paged KV pool as a contiguous cache)", e->B);
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
    if (e && e->B > 1) return fail(ZB_ESTATE, "this engine was created with batch %d: use the zb_engine_batch_* entry points (the single-sequence API would read the paged KV pool as a contiguous cache)", e->B);
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
    if (e && e->B > 1) return fail(ZB_ESTATE, "this engine was created with batch %d: use the zb_engine_batch_* entry points (the single-sequence API would read the paged KV pool as a contiguous cache)", e->B);
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
    CK(cudaMemcpy(host_out, e->final_hid ? e->final_hid : e->hid, (size_t)e->hidden * 4, cudaMemcpyDeviceToHost));
    return 0;
}

ZB_API int zb_engine_kv(zb_engine* e, int layer, int n, float* k_host, float* v_host) {
    if (!e || layer < 0 || layer >= e->layers || n < 0 || n > e->max_seq) return fail(ZB_EINVAL, "zb_engine_kv: bad arguments");
    CK(cudaSetDevice(e->opts.device));
    CK(cudaStreamSynchronize(e->stream));
    // device layout is [n_kv][max_seq][hd]; the tap returns rows [pos][n_kv*hd] like TensorCache.Get (tensor_cache.go:487-567)
    if (e->B > 1) return fail(ZB_ESTATE, "zb_engine_kv: a batch engine keeps its KV in a paged pool; this tap reads the single-sequence layout");
    const size_t hb = (size_t)e->hd * 4, pitch = (size_t)e->n_kv * hb;
    if (e->kv_f16) {   // fp16 cache: widen on the host
        std::vector<uint16_t> tmp((size_t)n * e->hd);
        for (int t = 0; t < 2; t++) {
            float* dst = t ? v_host : k_host;
            if (!dst) continue;
            const uint16_t* src = reinterpret_cast<const uint16_t*>(t ? e->L[layer].vc : e->L[layer].kc);
            for (int h = 0; h < e->n_kv && n > 0; h++) {
                CK(cudaMemcpy(tmp.data(), src + (size_t)h * e->max_seq * e->hd, tmp.size() * 2, cudaMemcpyDeviceToHost));
                for (int i = 0; i < n; i++)
                    for (int d = 0; d < e->hd; d++) dst[((size_t)i * e->n_kv + h) * e->hd + d] = __half2float(__ushort_as_half(tmp[(size_t)i * e->hd + d]));
            }
        }
        return 0;
    }
    for (int h = 0; h < e->n_kv && n > 0; h++) {
        if (k_host) CK(cudaMemcpy2D((char*)k_host + h * hb, pitch, e->L[layer].kc + (size_t)h * e->max_seq * e->hd, hb, hb, n, cudaMemcpyDeviceToHost));
        if (v_host) CK(cudaMemcpy2D((char*)v_host + h * hb, pitch, e->L[layer].vc + (size_t)h * e->max_seq * e->hd, hb, hb, n, cudaMemcpyDeviceToHost));
    }
    return 0;
}

ZB_API int zb_engine_batch_reset(zb_engine* e) {
    if (!e || e->B <= 1) return fail(ZB_EINVAL, "zb_engine_batch_reset: engine was not created with batch > 1");
    CK(cudaSetDevice(e->opts.device));
    e->free_blocks.clear();
    for (int i = e->pool_blocks - 1; i >= 0; i--) e->free_blocks.push_back(i);
    std::fill(e->h_btab.begin(), e->h_btab.end(), 0);
    std::fill(e->h_bpos.begin(), e->h_bpos.end(), 0);
    CK(cudaMemsetAsync(e->d_btab, 0, ((size_t)e->B * (e->max_blocks + 4 + e->n_kv) + 16) * sizeof(int), e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

ZB_API int zb_engine_batch_step(zb_engine* e, const int32_t* tokens, int32_t* next) {
    if (!e || e->B <= 1 || !tokens) return fail(ZB_EINVAL, "zb_engine_batch_step: bad arguments");
    CK(cudaSetDevice(e->opts.device));
    for (int b = 0; b < e->B; b++) {
        if (tokens[b] < 0 || tokens[b] >= e->vocab) return fail(ZB_EINVAL, "token ID %d out of range [0, %d)", tokens[b], e->vocab);
        e->h_bpin[b] = tokens[b];
    }
    CK(cudaMemcpyAsync(e->d_btok, e->h_bpin, (size_t)e->B * 4, cudaMemcpyHostToDevice, e->stream));
    if (int rc = batch_run_step(e)) return rc;
    CK(cudaMemcpyAsync(e->h_bpin + e->B, e->d_btok, (size_t)e->B * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (next) memcpy(next, e->h_bpin + e->B, (size_t)e->B * 4);
    return 0;
}

ZB_API int zb_engine_batch_decode_n(zb_engine* e, const int32_t* first_tokens, int n, int32_t* out_tokens, float* ms) {
    if (!e || e->B <= 1 || !first_tokens || n <= 0) return fail(ZB_EINVAL, "zb_engine_batch_decode_n: bad arguments");
    if (n > e->out_cap) return fail(ZB_EINVAL, "n=%d exceeds output capacity %d", n, e->out_cap);
    CK(cudaSetDevice(e->opts.device));
    if (e->h_bpos[0] + n > e->max_seq) return fail(ZB_ESTATE, "%d steps do not fit the KV cache (pos %d, capacity %d)", n, e->h_bpos[0], e->max_seq);
    for (int b = 0; b < e->B; b++) {
        if (first_tokens[b] < 0 || first_tokens[b] >= e->vocab) return fail(ZB_EINVAL, "token ID %d out of range", first_tokens[b]);
        e->h_bpin[b] = first_tokens[b];
    }
    CK(cudaMemcpyAsync(e->d_btok, e->h_bpin, (size_t)e->B * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(e->d_bnout, 0, 4, e->stream));
    // blocks for all n steps up front so the timed region is pure graph launches
    for (int i = 0; i < n; i++) {
        for (int b = 0; b < e->B; b++) e->h_bpos[b] += i;
        int rc = batch_ensure_blocks(e);
        for (int b = 0; b < e->B; b++) e->h_bpos[b] -= i;
        if (rc) return rc;
    }
    CK(cudaEventRecord(e->ev0, e->stream));
    for (int i = 0; i < n; i++) {
        if (e->graph_batch) CK(cudaGraphLaunch(e->graph_batch, e->stream));
        else { Counter cnt; if (int rc = enqueue_batch_step(e, cnt)) return rc; }
        for (int b = 0; b < e->B; b++) e->h_bpos[b]++;
    }
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (ms) CK(cudaEventElapsedTime(ms, e->ev0, e->ev1));
    if (out_tokens) CK(cudaMemcpy(out_tokens, e->d_bout, (size_t)n * e->B * 4, cudaMemcpyDeviceToHost));
    return 0;
}

ZB_API int zb_engine_batch_logits(zb_engine* e, float* host_out) {
    if (!e || e->B <= 1 || !host_out) return fail(ZB_EINVAL, "zb_engine_batch_logits: bad arguments");
    CK(cudaSetDevice(e->opts.device));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(host_out, e->b_logits, (size_t)e->B * e->vocab * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// Steady-state timing of one GEMV class for the roofline report: every streamed GEMV of block format `qtype` in the model
// (all layers, in model order, plain prologue, PDL-chained exactly as in the decode graph) is captured into one CUDA graph
// and replayed `reps` times between two events on the engine stream.  The class's weights of one step exceed the L2 by an
// order of magnitude, so every launch streams from HBM.  out->ms = event time of all reps.
ZB_API int zb_engine_profile_gemv_graph(zb_engine* e, int qtype, int reps, zb_gemv_profile* out) {
    if (!e || reps <= 0 || !out) return fail(ZB_EINVAL, "zb_engine_profile_gemv_graph: bad arguments");
    CK(cudaSetDevice(e->opts.device));
    std::vector<const DW*> ws;
    for (auto& L : e->L) {
        for (auto& w : L.qkv) ws.push_back(&w);
        ws.push_back(&L.o);
        for (auto& w : L.gate_up) ws.push_back(&w);
        if (L.down.main) ws.push_back(&L.down);
    }
    ws.push_back(&e->lm_head);
    int64_t maxk = 0, maxr = 0;
    std::vector<const DW*> sel;
    for (auto* w : ws)
        if (w->type == qtype && w->main && !w->e_main_stride) { sel.push_back(w); maxk = std::max(maxk, w->cols); maxr = std::max(maxr, w->rows); }
    if (sel.empty()) return fail(ZB_EINVAL, "no streamed GEMV of ggml type %d in this model", qtype);
    float *x = nullptr, *y = nullptr;
    if (int rc = dalloc(e, &x, (size_t)maxk * 2)) return rc;
    if (int rc = dalloc(e, &y, (size_t)maxr)) return rc;
    double bytes = 0;
    auto enqueue = [&]() -> int {
        for (auto* w : sel) {
            zb_prologue p{};
            p.a = x;
            p.eps = e->eps;
            DW plain = *w;
            plain.pairs = false;
            if (int rc = gemv(e, plain, p, y, e->use_pdl)) return rc;
        }
        return 0;
    };
    for (auto* w : sel) bytes += gemv_bytes(*w);
    if (int rc = enqueue()) return rc;  // warm (attributes, instruction cache)
    CK(cudaStreamSynchronize(e->stream));
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gx = nullptr;
    CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue();
    cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail((int)ce, "capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&gx, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail((int)ce, "instantiate failed: %s", cudaGetErrorString(ce));
    CK(cudaGraphLaunch(gx, e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    for (int i = 0; i < reps; i++) CK(cudaGraphLaunch(gx, e->stream));
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    cudaGraphExecDestroy(gx);
    out->qtype = qtype;
    out->launches = (int64_t)sel.size() * reps;
    out->bytes = bytes * reps;
    out->ms = ms;
    return 0;
}

// Tuning aid: phase stamps of the last persistent-kernel launch, [op][cta][8] SM-clock values (0 = not stamped), plus the
// op kinds.  Returns the number of ops (0 when the engine does not run the persistent kernel or ZB_MEGA_TRACE is unset).
ZB_API int zb_engine_tp_allreduce_us(zb_engine* e, int count, int reps, float* us, int* fused) {
    if (!e || count <= 0 || reps <= 0 || !us) return fail(ZB_EINVAL, "zb_engine_tp_allreduce_us: bad arguments");
    if (fused) *fused = e->tp_fused ? 1 : (e->tp_push ? 2 : 0);
    *us = 0.0f;
    if (e->tp_size <= 1) return 0;
    CK(cudaSetDevice(e->opts.device));
    float* buf = nullptr;
    if (int rc = dalloc(e, &buf, (size_t)e->hidden)) return rc;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gx = nullptr;
    Counter cnt;
    CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    e->ar_site = 0;
    e->ar_sites_per_step = std::max(count, 2 * e->layers);
    for (int i = 0; i < count && !rc; i++) rc = tp_allreduce(e, buf, (size_t)e->hidden, cnt);
    // every replay is a new "step" for the flagged exchange: its epochs must not match what the previous replay left behind
    if (!rc && e->d_step) step_bump_kernel<<<1, 1, 0, e->stream>>>(e->d_step);
    cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    if (rc || ce != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        return rc ? rc : fail((int)ce, "zb_engine_tp_allreduce_us: capture failed: %s", cudaGetErrorString(ce));
    }
    ce = cudaGraphInstantiate(&gx, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail((int)ce, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
    CK(cudaGraphLaunch(gx, e->stream));   // warm
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    for (int r = 0; r < reps; r++) CK(cudaGraphLaunch(gx, e->stream));
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    cudaGraphExecDestroy(gx);
    *us = ms * 1000.0f / (float)reps;
    return 0;
}

ZB_API int zb_engine_mega_trace(zb_engine* e, long long* out, int* kinds, int max_ops, int* ctas) {
    if (!e || !e->mega || !e->mctl.trace) return 0;
    cudaSetDevice(e->opts.device);
    cudaStreamSynchronize(e->stream);
    const int n = e->mctl.n_ops < max_ops ? e->mctl.n_ops : max_ops;
    if (ctas) *ctas = e->mega_ctas;
    if (out) cudaMemcpy(out, e->mctl.trace, (size_t)n * e->mega_ctas * kMegaTraceSlots * sizeof(long long), cudaMemcpyDeviceToHost);
    if (kinds) {
        std::vector<MegaOp> ops((size_t)n);
        cudaMemcpy(ops.data(), e->mctl.ops, (size_t)n * sizeof(MegaOp), cudaMemcpyDeviceToHost);
        for (int i = 0; i < n; i++) kinds[i] = ops[i].kind == kMegaGemv ? 100 + ops[i].g.type + 1000 * (ops[i].g.K / 256) : ops[i].kind;
    }
    return n;
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
