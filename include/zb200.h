/*
 * zb200.h -- B200-native entry points that sit beside the reference-compatible
 * kernel ABI (include/zerfoo_kernels.h) in the same libkernels.so.
 *
 * Two groups:
 *   1. zb_engine_*: the host-side decode engine.  It plays the role of the Go
 *      callers of the kernel library -- inference.LoadFile (inference/load_gguf.go:17-181),
 *      compute.WeightUploader.UploadWeights (load_gguf.go:101-116), generate.TensorCache
 *      (generate/tensor_cache.go:133-332), InferenceSession.Generate / graphForward
 *      (generate/session.go:84-268,440-461), runDecodeStep (generate/decode_step.go:26-67)
 *      and graph.NewCUDAGraphExecutor (generate/generator.go:301-365) -- as one C++
 *      object, because no Go toolchain exists in the build image (Boundary).
 *   2. zb_*: stand-alone launchers for B200-specific layouts and fused kernels that
 *      have no counterpart in the reference library.
 *
 * Conventions are the reference's: plain pointers and C ints, raw
 * device pointers, trailing cudaStream_t, cudaError_t-compatible int return
 * (0 = success); launchers are asynchronous, allocation-free and capture-safe.
 * Engine calls return 0 or a negative zb error / positive cudaError_t, and
 * zb_last_error() describes the most recent failure on the calling thread.
 */
#ifndef ZB200_H
#define ZB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* zb_stream_t; /* == cudaStream_t */
typedef struct zb_engine zb_engine;

/* ggml tensor type ids, as stored in GGUF (model/gguf/loader.go:140-190). */
enum { ZB_F32 = 0, ZB_F16 = 1, ZB_Q4_0 = 2, ZB_Q8_0 = 8, ZB_Q4_K = 12, ZB_Q5_K = 13, ZB_Q6_K = 14 };

typedef struct zb_engine_opts {
    int device;          /* CUDA device ordinal */
    int max_seq;         /* KV capacity; 0 = min(context_length, 4096) */
    int use_graph;       /* 1 = capture the decode step in a CUDA graph (ZERFOO_DISABLE_CUDA_GRAPH=1 overrides) */
    int tp_rank;         /* tensor-parallel rank / world (1 = single GPU) */
    int tp_size;
    int batch;           /* decode batch (sequences); 0/1 = single sequence */
    int flags;           /* ZB_ENGINE_* bits */
    int reserved[7];
} zb_engine_opts;
/* Decode-step variant for dense single-GPU batch-1 models whose matrices all have block-tiles.  The default is the
 * PDL-chained CUDA graph of per-matrix launches; ZB_ENGINE_MEGA (or ZB_MEGA=1 in the environment) selects the persistent
 * whole-token kernel (decode_mega.cu: one cooperative launch per token) instead -- measured slower on B200 so far
 * (profiles/r02), so it is opt-in.  ZB_ENGINE_NO_MEGA wins over both. */
#define ZB_ENGINE_NO_MEGA 1
#define ZB_ENGINE_MEGA 2
/* KV cache stored as fp16 (generate/tensor_cache.go:224-238 kvFP16 / WithKVDtype("fp16")): the append rounds K and V to
 * fp16, the decode attention reads half the bytes; arithmetic stays f32.  Also ZB_KV_F16=1 in the environment.
 * Chunked prefill and the persistent kernel keep the f32 cache (they refuse an fp16 engine). */
#define ZB_ENGINE_KV_F16 4

typedef struct zb_model_info {
    int vocab, hidden, layers, n_q, n_kv, head_dim, ffn, max_seq, n_experts, top_k;
    int tp_rank, tp_size;
    int64_t weight_bytes_per_token;  /* bytes of weights one decode step reads on this rank */
    int64_t kv_bytes_per_pos;        /* bytes of KV appended per position on this rank */
    int launches_per_step;           /* kernel launches in one captured decode step */
    char arch[32];
} zb_model_info;

const char* zb_last_error(void);

int zb_engine_create(const char* gguf_path, const zb_engine_opts* opts, zb_engine** out);
void zb_engine_destroy(zb_engine* e);
int zb_engine_info(const zb_engine* e, zb_model_info* out);

/* ---- tensor parallel (one process per GPU, NCCL over NVLink; inference/parallel/tensor_parallel.go:40-48) -----
 * QKV / gate / up / lm_head split by rows (heads, FFN columns, vocab), o / down split by block-aligned K columns,
 * one all-reduce of [hidden] after o_proj and after down_proj, experts sharded in contiguous blocks for MoE.
 * Rank 0 calls zb_tp_unique_id, the host broadcasts the 128 bytes, every rank calls zb_engine_create_tp. */
int zb_tp_unique_id(void* out128);
int zb_engine_create_tp(const char* gguf_path, const zb_engine_opts* opts, const void* nccl_id128, zb_engine** out);
/* host-side shard of a raw GGUF matrix: rows [r0, r1) x columns [c0, c1) (columns at block boundaries) */
int zb_tp_shard_host(int qtype, const void* raw, int64_t rows, int64_t cols, int64_t r0, int64_t r1, int64_t c0, int64_t c1, void* out);

/* cache.Reset + position counters to 0 (generate/session.go:118). */
int zb_engine_reset(zb_engine* e);

/* Runs the prompt through the stack (same per-token arithmetic as decode),
 * leaves logits of the last prompt token on the device, returns its greedy
 * argmax in *first_token (may be NULL). */
int zb_engine_prefill(zb_engine* e, const int32_t* tokens, int n, int32_t* first_token);

/* Prompt prefill in chunks of up to 256 tokens through the tcgen05 GEMMs (K-quant dense models, single sequence, no TP):
 * every matmul runs once per chunk over all its tokens, attention is causal over the cache.  bf16 operand rounding makes the
 * cache differ from the token-by-token prefill within the batched path's tolerance; zb_engine_prefill stays exact. */
int zb_engine_prefill_chunked(zb_engine* e, const int32_t* tokens, int n, int32_t* first_token, float* ms);

/* One decode step through the public path: H2D token, graph launch, D2H argmax
 * (generate/decode_step.go:26-67 + sampling_helpers.go:11-45). */
int zb_engine_decode_step(zb_engine* e, int32_t token, int32_t* next_token);

/* n chained decode steps with the token kept on the device (no host round trip);
 * the tokens produced are copied to out_tokens (host) once at the end.  If ms is
 * non-NULL it receives the CUDA-event time of the n steps on the engine stream. */
int zb_engine_decode_n(zb_engine* e, int32_t first_token, int n, int32_t* out_tokens, float* ms);

/* session.Generate with temperature 0: reset, prefill, n_new greedy tokens. */
int zb_engine_generate(zb_engine* e, const int32_t* prompt, int n_prompt, int n_new, int32_t* out_tokens);

/* Copies the logits of the last full step to host (vocab floats). */
int zb_engine_logits(zb_engine* e, float* host_out);
/* Debug/parity taps: hidden state after the last layer, and layer KV rows [0,n). */
int zb_engine_hidden(zb_engine* e, float* host_out);
int zb_engine_kv(zb_engine* e, int layer, int n, float* k_host, float* v_host);
int zb_engine_position(const zb_engine* e);
/* Tuning aid (ZB_MEGA_TRACE=1): SM-clock stamps of thread 0 of every CTA in the last persistent-kernel launch,
 * out[op][cta][8] = op start, fragments built, main loop done (warp 0), main loop done (CTA), op done, barrier passed;
 * kinds[op] = op kind (0 embed, 2 attention, 3 final) or 100 + ggml type + 1000 * (K / 256) for a GEMV.  Returns the op count. */
int zb_engine_mega_trace(zb_engine* e, long long* out, int* kinds, int max_ops, int* ctas);
zb_stream_t zb_engine_stream(const zb_engine* e);
/* Tensor-parallel engines: time `count` all-reduces of one hidden-size f32 vector (the exchange after o_proj / down_proj,
 * inference/parallel/tensor_parallel.go:151-163) as the decode step issues them -- back to back on the engine stream,
 * captured in a CUDA graph, `reps` replays between two events.  *us = microseconds per replay (= per decode step when
 * count = 2 * layers).  Collective: every rank must call it.  *fused: which exchange the engine uses -- 0 ncclAllReduce,
 * 2 the one-shot push all-reduce over peer memory (both timed here as the step issues them), 1 the exchange fused into the
 * GEMV epilogue / prologue (then the number is what the NCCL path WOULD cost). */
int zb_engine_tp_allreduce_us(zb_engine* e, int count, int reps, float* us, int* fused);

/* ---- batched decode (opts.batch > 1): `batch` sequences advance in lock-step over a paged KV cache
 * (16-position blocks from a shared pool, generate/paged_kv.go + block_pool.go) with tcgen05 GEMMs.
 * K-quant models only.  tokens / next are host arrays of `batch` ids. */
int zb_engine_batch_reset(zb_engine* e);
int zb_engine_batch_step(zb_engine* e, const int32_t* tokens, int32_t* next);
/* n chained steps, tokens stay on the device; out_tokens is [n][batch]; ms = CUDA-event time of the n steps */
int zb_engine_batch_decode_n(zb_engine* e, const int32_t* first_tokens, int n, int32_t* out_tokens, float* ms);
int zb_engine_batch_logits(zb_engine* e, float* host_out);   /* [batch][vocab] of the last step */

/* Per-format GEMV timing for the roofline report: `steps` eager decode steps with a
 * CUDA-event pair around every weight-streaming launch on the engine stream. */
typedef struct zb_gemv_profile {
    int qtype;
    int64_t launches;
    double bytes; /* algorithmic: weight blocks once + 4K + 4M per launch */
    double ms;    /* summed event time of those launches */
} zb_gemv_profile;
int zb_engine_profile_gemv(zb_engine* e, int steps, zb_gemv_profile* out, int max_classes, int* n_classes);
/* Steady-state variant: all GEMVs of one block format (every layer, model order, PDL-chained as in the decode graph) captured
 * into one CUDA graph and replayed `reps` times between two events; launches / algorithmic bytes / ms cover all reps. */
int zb_engine_profile_gemv_graph(zb_engine* e, int qtype, int reps, zb_gemv_profile* out);

/* ---- TMA-streamed fused GEMV (zerfoo_b200/csrc/gemv_stream.cu) --------------
 * The B200 replacement of Engine.MatMul on quantized storage at batch 1 plus the
 * fused providers around it (GPUFusedAddRMSNorm, GPUFusedNormAdd, FusedRMSNormGPU,
 * GPUFusedSwiGLU; SURVEY 8b "Go engine level").  Weights live in the stream
 * layout (16-B aligned rows + separate fp16 block scales), produced once at
 * upload time by zb_stream_repack_host -- the role UploadWeights plays for the
 * reference's separated Q4 layout (inference/load_gguf.go:101-116, gemm_q4.h:3-4). */
typedef struct zb_stream_weight {
    const void* main;        /* device: rows x row_main bytes */
    const void* aux;         /* device: fp16 block scales (Q4_0, Q8_0, Q6_K), padded by zb_stream_layout; else NULL */
    int qtype, rows, cols;
    /* MoE expert indirection (NULL / 0 for a plain matrix): slot k of n_sel uses expert expert_sel[k] */
    const int* expert_sel;   /* device */
    int n_sel, y_slot_stride;
    int64_t expert_main_stride, expert_aux_stride; /* bytes between experts */
    int epilogue;            /* 0: y[row] = dot.  1: rows (2i, 2i+1) are (gate_i, up_i): y[i] = silu(gate_i) * up_i (GPUFusedSwiGLU) */
    /* Fused tensor-parallel exchange (row-parallel o_proj / down_proj): instead of y, every output row is stored as an
     * 8-byte (value, epoch) pair into the partial-sum slot of THIS rank on every peer (NVLink peer memory, cudaIpc-mapped;
     * peer_out[p] points at uint2[rows]); epoch = epoch_base[0]*sites_per_step + site + 1.  The data is its own flag
     * (LL protocol): no fence, no ticket, no separate collective.  The consumer is zb_prologue.wait_*. */
    int n_peers, site, sites_per_step;
    float* peer_out[8];
    unsigned int* peer_flag[8];
    const int* epoch_base;   /* device: step counter, identical on all ranks */
    int* ticket;             /* device int, zero between launches */
} zb_stream_weight;

typedef struct zb_prologue {
    const float* a;          /* input vector [cols] (SwiGLU: [gate | up], 2*cols) */
    const float* r;          /* optional residual added after the first norm */
    const float* w1;         /* optional RMSNorm gain applied to a before the add */
    const float* w2;         /* optional RMSNorm gain applied after the add -> x */
    float* sum_out;          /* optional: the residual stream (a [*w1] + r), written once */
    const float* mix_w;      /* optional MoE combine: a = sum_k mix_w[k] * a[k*mix_stride + i] */
    int mix_n, mix_stride;
    float eps;
    int swiglu;              /* 1: x[i] = silu(a[i]) * a[cols + i] */
    int a_slot_stride;       /* with expert_sel: slot k reads a + k*a_slot_stride */
    /* consumer side of the fused exchange (n_wait > 0): a points at the local slots, uint2[mix_n][mix_stride]; every element
     * is polled until its epoch equals wait_epoch_base[0]*wait_sites_per_step + wait_site + 1, then summed in rank order */
    const unsigned int* wait_flags;
    const int* wait_epoch_base;
    int n_wait, wait_site, wait_sites_per_step;
    /* optional (zb_gemv_mma_f32): `a` exists as a_replicas identical copies a_replica_stride floats apart; CTA c reads copy
     * c % a_replicas -- spreads the L2 hot spot of a vector that all 148 SMs read at the same moment */
    int a_replicas, a_replica_stride;
} zb_prologue;

int zb_stream_layout(int qtype, int rows, int cols, int64_t* main_bytes, int64_t* aux_bytes);
int zb_stream_check(int qtype, int rows, int cols);
int zb_stream_repack_host(int qtype, const void* raw, int rows, int cols, void* main_out, void* aux_out);
/* flags bit 0: launch with programmatic stream serialization (PDL) so the weight
 * prefetch overlaps the previous kernel in the stream. */
int zb_gemv_stream_f32(const zb_stream_weight* w, const zb_prologue* p, float* y, int flags, zb_stream_t stream);

/* ---- tensor-core fused GEMV for batch-1 decode (zerfoo_b200/csrc/gemv_mma.cu) ----
 * Same contract as zb_gemv_stream_f32 (Engine.MatMul on Q4_K storage, gemv_q4k.cu:68-160, plus the fused prologue /
 * SwiGLU pair epilogue), but dequantisation and the contraction run on the tensor pipe (mma.sync f16, the nibbles as
 * exact fp16 subnormals, x as three fp16 terms, f32 accumulate).  Weights live in 16-row x 256-weight block-tiles
 * written by zb_mma_repack_host (pure byte permutation of the GGUF blocks).  `scratch` is a zero-initialised device
 * buffer of zb_mma_layout's scratch_bytes (row-tile tickets + partial sums of tiles shared by two CTAs); the kernel
 * leaves it zeroed/reusable, launches sharing it must be stream-ordered.  MoE combine / fused TP exchange prologues are
 * not supported here (cudaErrorInvalidValue): those launches stay on zb_gemv_stream_f32. */
typedef struct zb_mma_weight {
    const void* data;        /* device: block-tiles, zb_mma_layout's weight_bytes */
    int qtype, rows, cols;
    int epilogue;            /* as zb_stream_weight.epilogue */
    /* MoE expert indirection (NULL / 0 for a plain matrix), as in zb_stream_weight: `data` is a stack of experts of `rows`
     * rows each (rows % 16 == 0), expert_stride bytes apart (zb_mma_layout of one expert); slot k of n_sel multiplies expert
     * expert_sel[k] (device; < 0: not on this rank, the slot's output is zeroed) with a + k * zb_prologue.a_slot_stride and
     * writes y + k * y_slot_stride.  scratch must hold n_sel times zb_mma_layout's scratch_bytes. */
    const int* expert_sel;
    int n_sel, y_slot_stride;
    int64_t expert_stride;
} zb_mma_weight;
int zb_mma_check(int qtype, int rows, int cols);
int zb_mma_layout(int qtype, int rows, int cols, int64_t* weight_bytes, int64_t* scratch_bytes);
int zb_mma_repack_host(int qtype, const void* raw, int rows, int cols, void* out);
/* Work split of one launch over at most max_ctas CTAs (148, or 148 / n_sel with expert slots): out[12] = units per row, row tiles,
 * block-tiles, block-tiles per CTA, CTAs, block-tiles per warp, tiles per ring stage, ring stages, partial-sum slots per row tile,
 * row tiles per CTA (max), dynamic shared memory bytes, bytes per block-tile.  Host-side; no device access. */
int zb_mma_geometry(int qtype, int rows, int cols, int max_ctas, int* out);
int zb_gemv_mma_f32(const zb_mma_weight* w, const zb_prologue* p, float* y, void* scratch, int flags, zb_stream_t stream);
/* Tuning aid: with ZB_MMA_TRACE=1 the first 64 launches record 8 globaltimer stamps per CTA ([launch][148][8] uint64, ns);
 * returns the number of launches copied to `out`. */
int zb_mma_trace_read(unsigned long long* out, int max_launches);

/* ---- batched dequant-GEMM on tcgen05 / TMEM (zerfoo_b200/csrc/gemm_tc.cu) -----
 * Y[tokens, rows] = X[tokens, cols] . deq(W)^T for decode batches (>= 16 tokens) and prefill: the B200
 * replacement of gemm_q4_kernel N > 1 (gemm_q4.cu:116-159) and dequant_q4k_f32 + cuBLAS SGEMM
 * (dequant_q4k.cu:1-8).  K-quants (Q4_K, Q5_K, Q6_K) in the stream layout; weights are rounded once to
 * bf16 after the bit-exact f32 dequant, activations are bf16 hi (+ optional bf16 lo residual), f32 accumulate.
 * Activation operand ("tile image"): bf16 [cols/64][ldx][64] -- per 64-wide k-step a plane of ldx token rows (ldx % 16 == 0,
 * ldx >= tokens), each row's eight 16-byte chunks XOR-swizzled by (token & 7) and in the k-slot order of the format, so a
 * CTA's B tile is one contiguous bulk copy.  zb_gemm_tc_prep_x / zb_gemm_tc_prep_rows write it (ld_out / ldx = plane rows). */
int zb_gemm_tc_prep_x(int qtype, const float* x, int tokens, int K, int ldx, void* xhi, void* xlo, int ld_out, zb_stream_t stream);
/* Batched fused prologue (one row per token): the FusedAddRMSNorm / NormAdd / RMSNorm / SwiGLU providers of the reference
 * applied to [tokens, K] and written straight into the GEMM's bf16 operand (and optionally f32). */
typedef struct zb_prep_args {
    const float* a;          /* [tokens, lda] */
    const float* r;          /* optional residual [tokens, ldr] */
    const float* w1;         /* optional first RMSNorm gain [K] */
    const float* w2;         /* optional second RMSNorm gain [K] */
    float* sum_out;          /* optional residual stream out [tokens, ldsum] */
    float* x_f32;            /* optional f32 copy of x [tokens, ldxf] */
    void* xhi;               /* bf16 tile image [K/64][ldx][64] (see above), or NULL */
    void* xlo;
    int lda, ldr, ldsum, ldxf, ldx;
    float eps;
    int mode;                /* 0: norm/add chain; 1: SwiGLU over interleaved (gate_i, up_i) pairs in a[2K]; 2: SwiGLU over [gate | up] */
    int K, qtype;
} zb_prep_args;
int zb_gemm_tc_prep_rows(const zb_prep_args* a, int tokens, zb_stream_t stream);
int zb_gemm_tc_f32(const zb_stream_weight* w, const void* xhi, const void* xlo, int tokens, int ldx, float* y, int ldy, zb_stream_t stream);

/* ---- fused decode attention stage (zerfoo_b200/csrc/attention.cu) ------------
 * One launch per layer and token: per-head QK RMSNorm (optional) + half-split RoPE at
 * the device-resident position, KV append, split-KV flash decode over the
 * [n_kv][max_seq][head_dim] f32 cache, and the split merge.  head_dim in {32,64,128,256},
 * n_q/n_kv in {1,2,3,4,8}, chunk >= 16, max_splits*chunk >= max_seq.
 * Scratch: part_o [n_q*max_splits*head_dim], part_ml [2*n_q*max_splits], ticket [n_kv] ints (zeroed once). */
typedef struct zb_attn_args {
    const float* qkv;        /* [n_q*hd | n_kv*hd | n_kv*hd] projections of this token */
    const float* q_norm;     /* per-head RMSNorm gains [hd] or NULL */
    const float* k_norm;
    const float* cos_tbl;    /* [max_seq][hd/2] */
    const float* sin_tbl;
    const int* pos;          /* device: position of this token (kv_len = pos + 1) */
    void* k_cache;           /* f32, or fp16 when kv_f16 */
    void* v_cache;
    float* out;              /* [n_q*hd] */
    float* part_o;
    float* part_ml;
    int* ticket;
    float eps;
    int head_dim, n_q, n_kv, max_seq, chunk, max_splits;
    /* batched decode over a paged cache (0 / NULL = one sequence, contiguous cache): sequence b reads qkv + b*qkv_stride,
     * pos[b], block_table[b*max_blocks + t/page] into pools laid out [block][n_kv][page][head_dim]; scratch is per sequence. */
    int batch, qkv_stride, out_stride;
    const int* block_table;
    int max_blocks, page;
    int warps;               /* warps per CTA: 0 (default 4), 4, 8 or 16 -- 16 pays with long tiles (chunk >= 64) */
    /* Sliding window of the PROMPT pass (Mistral family): while *window_on != 0 the token at position p attends cache rows
     * (p - window, p] only -- the reference masks i - j >= window when the prompt goes through one Forward of seqLen > 1 and
     * attends the whole cache on decode steps (grouped_query_attention.go:1074-1077,1395-1415).  0 / NULL: no window. */
    int window;
    const int* window_on;
    int kv_f16;              /* 1: k_cache / v_cache hold fp16 (generate/tensor_cache.go:224-238 kvFP16): the append rounds, the
                              * tiles arrive as fp16 (half the KV bytes per step), scores and softmax stay f32 */
} zb_attn_args;
/* flags bit 0: PDL launch.  bit 1 (ZB_ATTN_SINGLE_TILE): the caller guarantees kv_len <= chunk * max_splits although
 * chunk * max_splits < max_seq (e.g. one long tile per KV head for short contexts: no split merge); a longer context traps. */
#define ZB_ATTN_SINGLE_TILE 2
int zb_decode_attn_f32(const zb_attn_args* a, int flags, zb_stream_t stream);

/* Chunked prefill attention for `tokens` prompt positions p0 .. p0+tokens-1 of one sequence: QK-norm + RoPE + KV append
 * into the [n_kv][max_seq][head_dim] cache, then causal attention of every query over cache rows [0, p0+i].
 * qkv: [tokens, ld_qkv] rows of (q heads | k heads | v heads); q_rot / out: [tokens, n_q*head_dim].
 * Default: q.k and p.v on the tensor cores with fp16 operands (f32 accumulate / softmax) for head_dim 32/64/128;
 * ZB_PREFILL_ATTN_F32 forces the all-f32 path (replaces flash_attention_forward_f32, flash_attention.cu:43-175). */
#define ZB_PREFILL_ATTN_F32 1
int zb_prefill_attn_f32(const float* qkv, int ld_qkv, const float* q_norm, const float* k_norm, const float* cos_tbl, const float* sin_tbl,
                        int p0, int tokens, float* q_rot, float* k_cache, float* v_cache, float* out, float eps, int head_dim, int n_q,
                        int n_kv, int max_seq, int window, int flags, zb_stream_t stream);

/* ---- stand-alone B200 launchers ------------------------------------------ */

/* Native GGUF Q8_0 (34 B blocks, fp16 scale) GEMV without the reference's 36 B repack. */
int zb_gemv_q8_0_f32(const void* W, const float* x, float* y, int M, int K, zb_stream_t stream);
/* Native GGUF Q4_0 (18 B interleaved blocks) GEMV. */
int zb_gemv_q4_0_f32(const void* W, const float* x, float* y, int M, int K, zb_stream_t stream);
/* Bit-exact dequantisation of n elements of any supported type to f32 (device pointers). */
int zb_dequant_f32(int qtype, const void* src, float* dst, int64_t n, zb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ZB200_H */
