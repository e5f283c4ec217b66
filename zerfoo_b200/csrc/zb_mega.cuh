// Persistent whole-token decode kernel ("megakernel"): program format shared by the device code (decode_mega.cu) and the
// host-side builder (engine.cu).  One cooperative launch of one CTA per SM (16 consumer warps + a TMA producer warp) walks
// a table of ops -- embedding gather, tensor-core GEMVs with their fused prologues / epilogues, decode-attention stages,
// lm_head + argmax + position bookkeeping.  The producer warp streams the block-tiles of ALL the CTA's GEMV ops through one
// ring of slots, so the next ops' weights arrive while the current op waits for its input, builds its activation fragments
// or stores its outputs.  Ops hand their vectors to each other as flagged (value, epoch) pairs (MegaVec below) instead of
// meeting at grid barriers; one grid barrier per launch remains, between the lm_head and the final argmax.
//
// Replaces, for dense single-GPU batch-1 decode, the PDL-chained CUDA graph of ~5 launches per layer (engine.cu:
// enqueue_step) -- and, in the reference, generate/megakernel.go:21-170 + internal/codegen/emit.go:131 (the generated
// whole-graph kernel, dead for GGUF graphs, generate/megakernel.go:28-32) and the CUDA-graph replay of
// generate/generator.go:301-365.  Opt-in (ZB_ENGINE_MEGA): measured slower than the graph step so far.
#pragma once
#include "zb_attn_tile.cuh"
#include "zb_mma_tiles.cuh"

namespace zb {

enum { kMegaEmbed = 0, kMegaGemv = 1, kMegaAttn = 2, kMegaFinal = 3 };
constexpr int kMegaThreads = (kMW + 1) * 32;   // 16 consumer warps + the TMA producer warp
constexpr int kMegaAttnWarps = 8;

// A vector one op of the launch hands to the next, in flagged form (the LL idea of collective libraries, applied to the
// activation vectors of a decode step): every element is an 8-byte (value bits, epoch) pair written with one store and polled
// by its readers until the epoch matches -- data and "ready" travel together, so GEMV -> GEMV needs no grid barrier and no
// fence: one store to L2 plus one poll instead of store, release, arrive, poll, reload.  A vector may come in several planes
// (partial sums of row tiles that straddle two CTAs): element i = plane 0 [i] + plane 1 [i] + ..., in that order.
// epoch = epoch_base (changes every launch) + the tag op's index, so stale pairs of earlier layers / tokens never match.
struct MegaVec {
    const uint2* p;
    int planes, stride;      // pairs between planes
    int tag_op;              // the pairs carry epoch_base + tag_op
    int pad;
};

struct MegaGemv {
    const uint8_t* w;        // block-tiles (zb_mma_repack_host)
    uint2* y;                // flagged output, y_planes planes of y_stride pairs
    float* y_plain;          // lm_head: plain logits instead
    MegaVec a, r;            // activation and (r.p != nullptr) residual
    const float* w1;         // Gemma-3 post-norm gain applied to a before the residual add
    const float* w2;         // RMSNorm gain of x
    uint2* sum_out;          // CTA 0 stores v = a' + r here, flagged with this op's tag
    float* sum_plain;        // ... and / or here as plain floats (the host-visible final hidden state)
    float eps;
    int swiglu;              // x[i] = silu(a[i]) * a[K + i]
    int tag_op;              // outputs carry epoch_base + tag_op (ops that fill one vector together share a tag)
    int y_planes, y_stride;
    int type, M, K, pairs;
    int nb, n_tiles, total, per_cta, per_warp, slots, max_local;      // work split = make_mgeom's (same summation order as gemv_mma_kernel)
    int xf_off, xm_off, xinv_off, part_off;                          // inside the CTA's scratch region
    int stream;              // index of this op's entry in the stream table
    int head;                // 1: lm_head -- skipped by launches without head, softcap + per-CTA argmax candidate in the epilogue
    float softcap;
};

struct MegaEmbed {
    const uint8_t* table;
    const int *feed, *feed_idx, *feed_len, *last;
    uint2* out;              // flagged with the embed op's own index
    int type, hidden, vocab;
    float scale;
};

struct MegaFinal {
    int *pos, *feed_idx, *amax, *last, *out, *n_out, *step;
    unsigned int* epoch_step;
    const int* feed_len;
    int out_cap;
};

struct MegaOp {
    int kind;
    int barrier;             // grid barrier after the op (the lm_head only: the final argmax reads every CTA's candidate)
    union {
        MegaGemv g;
        AttnArgs a;
        MegaEmbed e;
        MegaFinal f;
    };
};

// What a warp's TMA producer needs to know about GEMV op i, in op order (the weights are constants: the producer runs ahead
// of the consumer by as much as its ring holds, across op and barrier boundaries).
struct MegaStream {
    const uint8_t* w;
    int total, per_cta, per_warp, bt, pad0, pad1;
};

struct MegaCtl {
    const MegaOp* ops;
    const MegaStream* streams;
    int n_ops, n_streams, n_streams_nohead;
    int n_barriers;          // grid barriers per launch (identical with and without head)
    int region_bytes;        // per-CTA scratch: activation fragments / partial sums of a GEMV, or the K/V tile of an attention item
    int nslots, slot_bytes;  // the CTA's TMA ring: block-tile slots (after the scratch region), then full[nslots] / empty[nslots] mbarriers
    unsigned int* bar_counter;   // monotonic arrival counter of the grid barrier
    const int* step;         // launches since reset: barrier k of this launch completes at (step * n_barriers + k + 1) * gridDim.x
    const unsigned int* epoch_step;   // launches since creation (never reset): epoch_base = 1 + epoch_step * n_ops
    float* cand_v;           // per-CTA argmax candidates of the lm_head epilogue
    int* cand_i;
    long long* trace;        // optional phase timeline (ZB_MEGA_TRACE=1): [op][cta][8] SM-clock stamps of thread 0
};
constexpr int kMegaTraceSlots = 8;

// Dynamic shared memory of the launch.  192 KB + 1.8 KB static + 1 KB reserved fits the 196 KB carve-out, which leaves ~60 KB
// of L1: with the maximum carve-out (228 KB) every local-memory access and every __ldg is an L2 round trip.
constexpr int kMegaSmem = 192 * 1024;

// host side (decode_mega.cu)
int mega_max_ctas(int device, int* out_ctas);                 // co-resident CTAs of the kernel on this device (1 per SM)
int mega_launch(const MegaCtl& ctl, int ctas, int with_head, cudaStream_t stream);
bool mega_attn_supported(int head_dim, int rep);

}  // namespace zb
