/*
 * effort_b200.h -- C-ABI of the B200-native bucketMul hot path.
 *
 * Drop-in boundary for kolinko/effort's approximate GEMV.  The reference has no
 * FFI layer today: its boundary is a handful of Swift free functions plus the
 * deliberately non-private BucketMul.shared test hooks.  Each entry point below
 * names the reference interface (file:line under the reference checkout) that
 * it replaces.  All pointers named *_dev are CUDA device pointers; `stream` is
 * a cudaStream_t passed as void* (0 = legacy default stream).  Every call that
 * takes a stream only ENQUEUES work on it (the reference's gpu.deploy is
 * enqueue-only, helpers/gpu.swift:135-196; completion = gpu.eval(), :109-119).
 * All functions return 0 or a negative EFFORT_E* code; nothing aborts (the
 * reference asserts, bucketMul.swift:35-36).
 *
 * No torch / C++ types cross this boundary.  The library fails at load time if
 * the CUDA runtime is missing; there is no CPU fallback.
 */
#ifndef EFFORT_B200_H
#define EFFORT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFFORT_B200_VERSION 200

/* error codes */
#define EFFORT_OK 0
#define EFFORT_EINVAL (-1)    /* bad argument / failed reference precondition  */
#define EFFORT_ECUDA (-2)     /* CUDA runtime error (see effort_last_cuda_error) */
#define EFFORT_ENOMEM (-3)
#define EFFORT_ESHAPE (-4)    /* shape outside what the kernels support          */
#define EFFORT_ESTATE (-5)    /* e.g. effort_mul without a prior calc_dispatch    */
#define EFFORT_ENOTLOADED (-6)/* buckets not loaded and no dense core to fall back on */

/* weight kinds: main.swift:49-51 globals goQ4 / goQ8 (Q8 asserts false, expertMul.swift:33-36) */
#define EFFORT_KIND_FP16 0
#define EFFORT_KIND_Q4 1

/* effort_weights_create flags */
#define EFFORT_WEIGHTS_DEFAULT 0u
/* keep using the caller's rank-major buffers for the fast path too (no input-major device repack;
 * slower gathers, zero extra memory).  Without it the library builds its own repacked copy and the
 * caller's buckets/stats are only needed again by the effort_calc_dispatch/effort_mul test hooks. */
#define EFFORT_WEIGHTS_NO_REPACK 1u
/* device copy in the slice-major layout [column slice of 128][input][rank]: the rank rows an input selects inside a
 * column slice are contiguous, so a streaming unit is one cp.async.bulk (TMA) copy.  Round-2 engine only; the default
 * for FP16 weights (EFFORT_LAYOUT=input or EFFORT_ENGINE=1 in the environment makes input-major the default). */
#define EFFORT_WEIGHTS_SLICE_MAJOR 2u
/* device copy input-major ([input][rank][all columns]) even when slice-major is the default: what the round-1 engine reads */
#define EFFORT_WEIGHTS_INPUT_MAJOR 4u

#define EFFORT_PROBES_COUNT 4096 /* bucketMul.swift:19, loader.swift:66 */

typedef struct effort_ctx effort_ctx_t;         /* replaces the BucketMul.shared / BucketMulQ4.shared singletons
                                                   (bucketMul.swift:18-32): per-context scratch, one stream at a time */
typedef struct effort_weights effort_weights_t; /* replaces class ExpertWeights (loader.swift:46-167) */

/* ---- library ---------------------------------------------------------- */
int effort_version(void);
const char* effort_strerror(int code);
/* text of the last CUDA error seen by this thread's calls ("" if none) */
const char* effort_last_cuda_error(void);

/* ---- context ---------------------------------------------------------- */
/* Scratch: dispatch list (maxDispatchSize = 229376*2 float2 entries, bucketMul.swift:20-29),
 * cutoff scalar, partial-sum buffer (tmpMulVec [32,16384], bucketMul.swift:52 -> here [nCTA, out]).
 * `device` = CUDA device ordinal (-1 = current). */
int effort_ctx_create(int device, effort_ctx_t** ctx_out);
int effort_ctx_destroy(effort_ctx_t* ctx);
/*
 * How calcDispatch's cutoff (findCutoff32, bucketMul.metal:141-247) is computed by the fused operator:
 *   EFFORT_CUTOFF_SELECT (default)  the exact order statistic the reference's bisection approximates: the (k+1)-th
 *                                   largest of the 4096 bf16 probe products, k = 4096 - Int(4095*(1-effort)), so exactly
 *                                   k products lie above it (0 when k = 4096) -- a radix select, ~1 us;
 *   EFFORT_CUTOFF_BISECT            the reference's loop replayed bit for bit (same fp32 result and loop count).
 * Both satisfy the reference's own acceptance rule (count within 2 of k, bucketMul.metal:236).  effort_find_cutoff and
 * effort_calc_dispatch (the test hooks) always run the bisection.  Environment default: EFFORT_CUTOFF=bisect.
 */
#define EFFORT_CUTOFF_SELECT 0
#define EFFORT_CUTOFF_BISECT 1
int effort_ctx_set_cutoff_mode(effort_ctx_t* ctx, int mode);
/* Tuning / A-B knobs of the fused operator (tests and tools; every value computes the same operator):
 *   "engine"   2 (default) round-2 kernel: one launch per group, staged streaming, reductions into `out`;
 *              1 round-1 kernel + integrate launch (deterministic fp32 order)
 *   "stage"    4 (default) eight consumer warps accumulate, eight producer warps stage whole-input units (1..16 rows)
 *              with bulk async copies (cp.async.bulk, several units per producer step) into per-pair rings and hand them over
 *              through mbarriers (slice-major FP16 weights; other weights take stage 0); 3 the same pairs fed by 16-byte
 *              cp.async, one unit per producer step; 2 one TMA producer warp, a shared byte ring and 16 consumers (measured
 *              slower: the single producer's serial issue is the limit); 0 sixteen self-serving warps with private cp.async
 *              rings, units of at most 4 rows
 *   "window"   1..8 (default 8) stage 4: most units a producer takes per ticket grab
 *   "lookahead" 1 (default) / 0 stages 3-4: consumers test the next slot's barrier and fetch its descriptor early
 *   "dynamic"  per-warp rings (stage 0/1) only: 0 (default) static round robin of the units, 1 units from a shared counter
 *   "hint"     1 (default) stages 3-4: the exact select starts its search at the cutoff the same matrix produced on the
 *              previous call (3 rounds instead of 8 when it moved by less than 12 %; the result never depends on it)
 *   "prefetch" 0 (default) / 1 stages 3-4: while the cutoff is being computed, rows that the matrix's previous cutoff
 *              would select are prefetched into L2 (measured: no gain -- the gather is not DRAM-latency bound)
 * Returns EFFORT_EINVAL for an unknown name or value.  Environment defaults: EFFORT_ENGINE, EFFORT_STAGE (ldgsts|tma|pairs-ldgsts), EFFORT_WINDOW, EFFORT_LOOKAHEAD,
 * EFFORT_DYN, EFFORT_PREFETCH, EFFORT_HINT. */
int effort_ctx_set_option(effort_ctx_t* ctx, const char* name, int value);
/* Non-zero once a kernel of this context gave up a bounded wait: its output is invalid.  1 = the overwrite protocol
 * of a fused GEMV (co-resident CTAs never arrived), 2 / 3 = a consumer / producer warp of bucket_mul_v4 waited ~1 s on
 * its ring, 4 = a tensor-parallel exchange waited 2 s for a peer's packets (peer died or launched in another order).
 * Synchronises `stream`. */
int effort_ctx_error_flag(effort_ctx_t* ctx, unsigned* flag_out, void* stream);

/* ---- weights ---------------------------------------------------------- */
/*
 * ExpertWeights (loader.swift:46-167).  Shapes as the reference stores them:
 *   FP16: buckets [n_experts, in*percent_load, out/16] f16    stats [n_experts, in*percent_load, 4] f16
 *   Q4:   buckets [n_experts, in*8, out/32] u16 (viewed f16)  stats [n_experts, in*8, 2] f32
 *   probes [n_experts, 4096] f16;  outliers [n_outliers, 4] f32 or NULL (Q4);
 *   core [out, in] f16 or NULL (dense fallback, expertMul.swift:29-31).
 * percent_load = number of rank rows kept per input dim (16 = all, loader.swift:50,157-159); Q4 uses 8.
 * buckets_dev may be NULL only when core_dev is given ("buckets not loaded", loader.swift:105-108).
 * The handle never owns the caller's buffers.
 */
int effort_weights_create(const void* buckets_dev, const void* stats_dev, const void* probes_dev,
                          const void* outliers_dev, int n_outliers, const void* core_dev,
                          int in_dim, int out_dim, int n_experts, int percent_load, int kind,
                          unsigned flags, void* stream, effort_weights_t** w_out);
int effort_weights_destroy(effort_weights_t* w);
/* bytes of device memory the handle owns (the repacked copy) */
size_t effort_weights_owned_bytes(const effort_weights_t* w);

/* ---- the operator ------------------------------------------------------ */
/*
 * bucketMul(v:by:expNo:out:effort:)  bucketMul.swift:11  (FP16; overwrites out, bucketMul.metal:133)
 *   v_dev [in] f32, out_dev [out] f32, exp_no_dev: device uint32 (NULL = expert 0,
 *   expertMul.swift:18), effort in [0,1] (default 0.25 in the reference).
 */
int effort_bucket_mul(effort_ctx_t* ctx, const float* v_dev, const effort_weights_t* w,
                      const uint32_t* exp_no_dev, float* out_dev, double effort, void* stream);
/* bucketMulQ4(v:by:expNo:out:effort:)  bucketMulQ4.swift:11  (accumulates INTO out, then calcOutliers) */
int effort_bucket_mul_q4(effort_ctx_t* ctx, const float* v_dev, const effort_weights_t* w,
                         const uint32_t* exp_no_dev, float* out_dev, double effort, void* stream);
/* expertMul(v:by:expNo:out:effort:)  expertMul.swift:24-38: Q4 -> out.zero(); bucketMulQ4 if buckets
 * loaded else basicMul(core); FP16 -> bucketMul. */
int effort_expert_mul(effort_ctx_t* ctx, const float* v_dev, const effort_weights_t* w,
                      const uint32_t* exp_no_dev, float* out_dev, double effort, void* stream);
/* basicMul(v:by:out:)  helpers/mps.swift:14-47: dense fp16 [out,in] x fp16(v) -> fp32 out */
int effort_basic_mul(effort_ctx_t* ctx, const float* v_dev, const void* core_dev, int out_dim, int in_dim,
                     float* out_dev, void* stream);

/*
 * Several bucketMuls that do not depend on each other (q/k/v share v, runNetwork.swift:132-134;
 * w1/w3, :178-179) enqueued as ONE launch group.  Semantically identical to n calls of
 * effort_expert_mul in order.
 */
typedef struct {
    const float* v_dev;
    const effort_weights_t* w;
    const uint32_t* exp_no_dev;
    float* out_dev;
    double effort;
    /* Tensor-parallel row shards only: the first 4096 entries of the FULL input vector,
     * which findCutoff32 probes (bucketMul.metal:158-163), when v_dev is a rank-local slice.  NULL = v_dev. */
    const float* v_cutoff_dev;
} effort_mul_args_t;
int effort_expert_mul_batch(effort_ctx_t* ctx, const effort_mul_args_t* args, int n, void* stream);

/* ---- test hooks (kept callable "for testing", bucketMul.swift:17) ------- */
/* BucketMul.calcDispatch  bucketMul.swift:34-48: findCutoff32 + prepareDispatch into ctx scratch
 * (reference dispatch format: float2 {v, float(rowOffset)}; here in ascending row order), then the
 * roundUp/zeroRange32 padding of fullMul (bucketMul.swift:57-58). Works for both kinds. */
int effort_calc_dispatch(effort_ctx_t* ctx, const float* v_dev, const effort_weights_t* w,
                         const uint32_t* exp_no_dev, double effort, void* stream);
/* BucketMul.mul  bucketMul.swift:69-88: MAC over the ctx dispatch list + integrate (FP16: overwrites
 * out; Q4: accumulates into out). */
int effort_mul(effort_ctx_t* ctx, const effort_weights_t* w, float* out_dev, void* stream);
/* findCutoff32 only (bucketMul.metal:141-247); result stays in ctx scratch */
int effort_find_cutoff(effort_ctx_t* ctx, const float* v_dev, const effort_weights_t* w,
                       const uint32_t* exp_no_dev, double effort, void* stream);
/* Synchronising reads of ctx scratch (host pointers).  n_selected = rows selected before padding,
 * padded_size = dispatch.size after roundUp. dispatch_host may be NULL; else it must hold
 * 2*capacity floats and receives min(padded_size, capacity) entries. */
int effort_read_dispatch(effort_ctx_t* ctx, float* dispatch_host, size_t capacity, uint32_t* n_selected,
                         uint32_t* padded_size, float* cutoff, int* cutoff_loops, void* stream);

/* ---- convert ------------------------------------------------------------ */
/*
 * bucketize(_:outTensorsPref:tensors:)  convert.swift:209-260 (goQ8 = false): w_dev [out,in] f16 ->
 * buckets [in*16, out/16] f16, stats [in*16, 4] f16, probes [4096] f16, byte-identical to the
 * reference layout.  Preconditions of convert.swift:210-215 are returned as EFFORT_EINVAL.
 */
int effort_bucketize(const void* w_dev, int out_dim, int in_dim, void* buckets_dev, void* stats_dev,
                     void* probes_dev, void* stream);
/*
 * Q4 convert(core2)  q4_draft.py:70-322 after outlier extraction: wT_dev is W^T [in,out] f16 with the
 * outliers already zeroed (outlier selection is a global top-2% sort done once on the host side, see
 * effort_b200/convert.py) -> buckets [in*8, out/32] u16, stats [in*8, 2] f32, probes [min(in,out)] f16.
 */
int effort_q4_bucketize(const void* wT_dev, int in_dim, int out_dim, void* buckets_dev, void* stats_dev,
                        void* probes_dev, void* stream);

/* ---- tensor-parallel plumbing ------------------------------------------------- */
/*
 * One process per GPU.  The collectives of the sharded decode loop (all-gather of the row-parallel GEMV's
 * cutoff input, all-reduce of its partial output) are NCCL calls enqueued on the caller's stream from inside
 * the library (so they can live in the token's CUDA graph).  libnccl is resolved at run time (the copy already
 * loaded by the host process, else dlopen("libnccl.so.2")).  Rank 0 creates the 128-byte id and hands it to the
 * other ranks through whatever transport the host has (a process-group broadcast in effort_b200/model.py).
 */
int effort_comm_unique_id(void* id128_out);
int effort_comm_init(effort_ctx_t* ctx, const void* id128, int rank, int world);
int effort_comm_destroy(effort_ctx_t* ctx);
/* One-shot NVLink collectives (csrc/comm.cuh): every rank maps every peer's symmetric buffer through CUDA IPC and a
 * collective is one kernel per rank (peer stores + release flag + acquire spin), CUDA-graph replayable.  Used by the
 * sharded decode loop when connected (EFFORT_P2P=0 forces NCCL).  local_handle: allocates this rank's buffer and
 * returns its 64-byte cudaIpcMemHandle_t; connect: takes all ranks' handles (world x 64 bytes, rank order). */
int effort_comm_p2p_local_handle(effort_ctx_t* ctx, void* handle64_out);
int effort_comm_p2p_connect(effort_ctx_t* ctx, const void* handles, int rank, int world);
/* Test hook for the one-shot NVLink collectives themselves (the decode loop calls them internally): mode 0 = all-gather
 * (send: count floats, out: world*count floats in rank order), mode 1 = all-reduce (out: count floats, summed in rank
 * order); `site` (0..15) names the call position -- every rank must use the same sequence of sites, and a site may be
 * reused only after a later site's collective.  EFFORT_ESTATE unless effort_comm_p2p_connect succeeded. */
int effort_comm_p2p_collective(effort_ctx_t* ctx, int mode, int site, const float* send_dev, float* out_dev, size_t count,
                               void* stream);
/* back to NCCL for the exchanges (call on EVERY rank when any rank failed to connect: the choice must be collective) */
int effort_comm_p2p_disable(effort_ctx_t* ctx);
/* in-place sum all-reduce / all-gather of fp32 device buffers over the ctx communicator (test + building block) */
int effort_comm_all_reduce(effort_ctx_t* ctx, float* buf_dev, size_t count, void* stream);
int effort_comm_all_gather(effort_ctx_t* ctx, const float* send_dev, float* recv_dev, size_t send_count, void* stream);

/* ---- decode loop (host orchestration of the callers either side of the path) ----------------------- */
/*
 * runNetwork(tokens:effort:)  runNetwork.swift:68-316, one token per call: per layer rmsNormFast*attnNorm,
 * expertMul x3 (wq,wk,wv), rope_mx + calcScores + softmax + sumScores, expertMul wo, residual, rmsNormFast*
 * ffnNorm, expertMul w1,w3, silu, expertMul w2, residual (:124-183); final rmsNorm*norm and the dense
 * lm_head basicMul (:206-209); greedy next token = top-1 (:235-257).  north_star keeps this orchestration in
 * Swift; there is no Swift toolchain here, so the mirror lives behind the same C-ABI and a Swift build would
 * call the per-operator entry points above instead.  Dims follow main.swift:45-46,56,72-77.
 */
typedef struct effort_model effort_model_t;
typedef struct {
    int dim;         /* stateDim 4096 */
    int hidden_dim;  /* hiddenDim 14336 */
    int n_layers;    /* numLayers 32 */
    int n_heads;     /* numHeads 32 */
    int n_kv_heads;  /* 8 (kvRepeats = 4) */
    int head_dim;    /* headDim 128 (the attention kernel requires 128) */
    int vocab;       /* 32000 */
    int max_seq;     /* maxSeqLen 2048 */
    float rope_theta;/* 1e6: freqs = 1e-6^(j/64), model.swift:701 */
    float norm_eps;  /* 1e-5, aux.metal:151 */
    int tp_rank, tp_size; /* tensor-parallel shard of this process (tp_size 1 = unsharded).  With tp_size = G > 1 the
                             ctx must have a communicator (effort_comm_init) and the weights passed to
                             effort_model_set_layer / set_head are this rank's shards: wq [dim -> dim/G],
                             wk/wv [dim -> kv_dim/G], wo [dim/G -> dim] (row shard), w1/w3 [dim -> hidden/G],
                             w2 [hidden/G -> dim] (row shard), output.core [vocab/G, dim]. */
} effort_model_config_t;

int effort_model_create(effort_ctx_t* ctx, const effort_model_config_t* cfg, effort_model_t** m_out);
int effort_model_destroy(effort_model_t* m);
/* Layer weights (loader.swift:201-225).  The handles and norm vectors (fp16 [dim]) stay owned by the caller. */
int effort_model_set_layer(effort_model_t* m, int layer, const effort_weights_t* wq, const effort_weights_t* wk,
                           const effort_weights_t* wv, const effort_weights_t* wo, const effort_weights_t* w1,
                           const effort_weights_t* w2, const effort_weights_t* w3, const void* attn_norm_dev,
                           const void* ffn_norm_dev);
/* Mixture of experts (runNetwork.swift:185-200, loader.swift:208-212): `gate_dev` = layers.N.feed_forward.gate [n_experts,
 * dim] f16; the layer's w1 / w2 / w3 handles must hold n_experts experts.  Per token: gate logits = basicMul(rmsNorm(h) *
 * ffn_norm, gate), the two largest (device indices, read by the expert GEMVs as expNo -- no host sync), softmax over
 * the two, h += gateVal_i * w2_e(silu(w1_e x) * w3_e x) for both.  Fused chain only (effort_model_set_chain 2). */
int effort_model_set_moe(effort_model_t* m, int layer, const void* gate_dev, int n_experts);
/* model.norm [dim] f16, output.core [vocab,dim] f16, tok_embeddings.core [vocab,dim] f16 (loader.swift:254-272) */
int effort_model_set_head(effort_model_t* m, const void* norm_dev, const void* output_core_dev,
                          const void* tok_embeddings_dev);
/* position <- 0 (the KV cache is logically emptied) */
int effort_model_reset(effort_model_t* m, void* stream);
/*
 * One decode step at the current position, enqueue-only.  token_dev: device int32 (NULL = the token the
 * previous step predicted).  After it: logits in effort_model_logits(), next token in effort_model_next_token().
 * The first call for a given effort value runs eagerly and captures a CUDA graph; later calls replay it.
 */
int effort_model_step(effort_model_t* m, const int32_t* token_dev, double effort, void* stream);
/* Same step driven with HOST buffers (the end-to-end call): token_host (NULL = previous prediction) is copied
 * H2D from pinned memory, the step runs, next token (and logits if logits_host != NULL) are copied D2H and the
 * stream is synchronised. */
int effort_model_step_host(effort_model_t* m, const int32_t* token_host, double effort, int32_t* next_token_host,
                           float* logits_host, void* stream);
const float* effort_model_logits(const effort_model_t* m);        /* device [vocab] f32 */
const int32_t* effort_model_next_token(const effort_model_t* m);  /* device int32 */
/* bytes of bucket weights one token streams at effort 1.0 (sum of in*out*2 over the 7x n_layers matrices) */
size_t effort_model_bucket_bytes(const effort_model_t* m);
/* use CUDA graphs for effort_model_step (default 1) */
int effort_model_set_graphs(effort_model_t* m, int enable);
/* round-1 engine only (default 0): apply rmsNorm*w on load inside the round-1 bucketMul kernels and fold the residual
 * add and silu*mul into the integrate epilogues (9 launches per layer instead of 12) */
int effort_model_set_fused_glue(effort_model_t* m, int enable);
/* Token chain: 2 (default; single GPU, FP16 buckets, round-2 engine) = 5 launches per layer -- [q,k,v] with
 * rmsNorm*w applied on load, attention, wo accumulating into the residual stream, [w1,w3] with rmsNorm*w on load, w2
 * with silu(x1)*x3 on load accumulating into the residual stream -- then one head kernel (final norm + lm_head +
 * argmax); 1 = one kernel per reference op (runNetwork.swift:124-183 order).  Other configurations (tensor parallel,
 * Q4, round-1 engine) always use chain 1.  Environment default: EFFORT_CHAIN. */
int effort_model_set_chain(effort_model_t* m, int chain);

/* ---- introspection used by bench / tests -------------------------------- */
/* number of kernels this library has launched since load (process-wide) */
uint64_t effort_launch_count(void);
/* rows selected by the last fused bucket_mul on this ctx (synchronises `stream`) */
int effort_last_selected(effort_ctx_t* ctx, uint32_t* n_selected, void* stream);

/* ---- bucketed-safetensors model directory (host side, no CUDA) --------------------------------------------
 * Replaces TensorLoader (helpers/safetensors.swift:87-216): `<model>.safetensors.index.json` maps tensor names to
 * the per-layer files written by TensorSaver.save (:38-85); a name missing from the index is retried as
 * name + ".weight" (:141-146); only BF16 / F16 / F32 tensors are accepted (:176); data_offsets must span
 * prod(shape) * sizeof(dtype) bytes (:182).  Files are mapped read-only and stay mapped until effort_loader_close,
 * so `data` can be passed to cudaMemcpy (or wrapped without a copy).  ExpertWeights (loader.swift:113-166) keeps
 * only the first percentLoad * inDim rows of `buckets` / `bucket.stats`: rows are the leading dimension, so that
 * truncation is a prefix of `data` (see effort_b200/weights_io.py, NativeTensorLoader.expert_weights). */
#define EFFORT_ST_F16 0
#define EFFORT_ST_BF16 1
#define EFFORT_ST_F32 2
#define EFFORT_ST_MAX_DIMS 8
typedef struct effort_loader effort_loader_t;
typedef struct {
    int dtype;                         /* EFFORT_ST_* */
    int ndim;
    int64_t shape[EFFORT_ST_MAX_DIMS];
    const void* data;                  /* inside the read-only mapping */
    size_t nbytes;
} effort_tensor_info_t;
int effort_loader_open(const char* dir, const char* model, effort_loader_t** out);   /* TensorLoader(path:model:) :105-110 */
void effort_loader_close(effort_loader_t* loader);
int effort_loader_count(const effort_loader_t* loader);
const char* effort_loader_name(const effort_loader_t* loader, int i);                /* index order */
int effort_loader_has(const effort_loader_t* loader, const char* name);              /* hasTensor :132-134 */
int effort_loader_tensor(effort_loader_t* loader, const char* name, effort_tensor_info_t* info);  /* fetchTensor :136-216 */
/* convertBF16 (the reference converts BF16 tensors to fp16 after loading, safetensors.swift:207-210) */
int effort_bf16_to_f16(const uint16_t* src, uint16_t* dst, size_t n);

/* Replaces TensorSaver (helpers/safetensors.swift:38-85) + saveSafetensors (:222-280): tensors are collected per output
 * file (the reference writes one file per layer, convert.swift:68,121) and effort_saver_save writes
 * "<model>-%05d-of-%05d.safetensors" files (u64 LE header size, JSON header with dtype / shape / data_offsets per tensor and
 * "__metadata__": {"description"}, tensor bytes in header order) plus "<model>.safetensors.index.json"
 * {"weight_map": {name: file}}.  Only F16 / F32 (and BF16 pass-through) tensors, as in the reference (:231-247).  The saver
 * does not copy: `data_host` must stay valid until effort_saver_save returns.  A name may be added once (EFFORT_ESTATE). */
typedef struct effort_saver effort_saver_t;
int effort_saver_open(const char* dir, const char* model, const char* description /* NULL = the reference's text */,
                      effort_saver_t** out);
int effort_saver_add(effort_saver_t* saver, int file_index, const char* name, int dtype, int ndim, const int64_t* shape,
                     const void* data_host, size_t nbytes);
int effort_saver_save(effort_saver_t* saver);
void effort_saver_close(effort_saver_t* saver);

#ifdef __cplusplus
}
#endif
#endif /* EFFORT_B200_H */
