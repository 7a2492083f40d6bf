/*
 * tethys.h — C-ABI of libtethys.so: the B200-native (sm_100a) replacement for the TensorFlow runtime
 * slice that tethys-speech's data-parallel train step executes.
 *
 * The reference (hyunnnchoi/tethys-speech) has no FFI of its own: its step is Python calling TF ops.
 * Every entry point below therefore cites the reference *call site* whose TF op(s) it replaces
 * (W = speech_jobs/whisper_dist.py, V = speech_jobs/wav2vec2_dist.py, VS = wav2vec2_single.py,
 * WS = whisper_single.py).  The Python host (the modules under tethys_speech_b200/) binds these with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless named host_*;
 *   - the library borrows inputs and writes into caller-allocated outputs; it owns nothing but the
 *     ts_ctx and the model-program handles;
 *   - every call is asynchronous on the caller-supplied cudaStream_t (passed as void*);
 *   - return 0 on success, negative ts_status on failure; message via ts_last_error();
 *   - there is NO CPU fallback: unsupported shapes/dtypes return TS_EUNSUPPORTED.
 */
#ifndef TETHYS_H_
#define TETHYS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TS_VERSION 100 /* 0.1.0 */

typedef enum {
  TS_OK = 0,
  TS_EINVAL = -1,
  TS_ESHAPE = -2,
  TS_EDTYPE = -3,
  TS_ECUDA = -4,
  TS_ENCCL = -5,
  TS_EUNSUPPORTED = -6,
  TS_EWATCHDOG = -7 /* a device-side mbarrier wait timed out (pipeline bug), see ts_watchdog_check */
} ts_status;

typedef enum { TS_F32 = 0, TS_BF16 = 1, TS_I32 = 2, TS_I64 = 3 } ts_dtype;

typedef struct ts_ctx ts_ctx;

/* ---- context ------------------------------------------------------------------------------- */
int ts_version(void);
/* one context per (process, GPU); replaces TF's per-device runtime state. */
int ts_create(int device, ts_ctx** out);
void ts_destroy(ts_ctx* ctx);
const char* ts_last_error(ts_ctx* ctx);
/* returns TS_EWATCHDOG if any kernel since the last check gave up on an mbarrier wait (synchronises). */
int ts_watchdog_check(ts_ctx* ctx);
/* number of kernel-launch sites this context has passed since creation (bench.py's gpu_launches). */
int64_t ts_launch_count(ts_ctx* ctx);
/* number of bf16 GEMMs that could not be expressed as TMA tiles (unaligned pointer / leading dimension) and ran on the fp32
 * CUDA-core engine instead of tcgen05 since creation: a perf cliff, never a numerical one. bench.py reports it; with the
 * environment variable TETHYS_STRICT_TC=1 such a GEMM returns TS_EUNSUPPORTED instead. */
int64_t ts_simt_downgrades(ts_ctx* ctx);
/* debug: `device_buf` (>= 16 * grid uint64, or NULL to stop) receives up to 16 %globaltimer stamps per CTA from every following tcgen05
 * GEMM launch: 0 entry, 1 setup done, 2 first operands landed, 3 first tile issued, 4 last MMA issued, 5/6 accumulator ready
 * (a tile / the last tile), 7 stores drained. tools/gemm_trace.py turns them into a head / main loop / tail breakdown. */
int ts_debug_gemm_trace(ts_ctx* ctx, void* device_buf);

/* ---- device-resident step state (CUDA-graph replay of a whole train step) ----------------------------------------------
 * The library keeps {dropout salt, optimizer step} in device memory. Every dropout kernel adds the salt to its seed, and
 * ts_optim_step(step = 0) takes the Adam step count from there. A captured step that starts with ts_step_state_advance
 * therefore draws fresh dropout masks and uses the right bias correction on every replay, with no host involvement.
 * Default state is {0, 0}: eager callers that pass their own seeds / step numbers are unaffected.
 */
int ts_step_state_set(ts_ctx* ctx, uint64_t salt, int64_t step, void* stream);
int ts_step_state_advance(ts_ctx* ctx, void* stream);
/* reads the state back (synchronises the device): checkpoints store it so that a restored run draws the masks the straight
 * run would have drawn, also when the step is replayed from a CUDA graph. */
int ts_step_state_get(ts_ctx* ctx, uint64_t* salt, int64_t* step);

/* ---- K9: GEMM with fused epilogue ------------------------------------------------------------
 * Replaces every tf.keras.layers.Dense / tf.matmul / Conv1D-as-GEMM on the path:
 *   W:89-92,141,174,194-205,311-312,545 ; V:240-268,316-319,338-340,371,383-398,553,579,586 and
 *   their autodiff transposes (dgrad, wgrad).
 *
 *   C[m,n] = dropout(act(alpha * sum_k A[m,k]*B[n,k] + bias[n])) + residual[m,n]   (per batch)
 *
 * Storage ("major") of the two operands, in elements:
 *   a_major = 0 ("K-major"):  A stored as [m][k], row stride lda   (reduce dim contiguous)
 *   a_major = 1 ("MN-major"): A stored as [k][m], row stride lda   (m contiguous)
 *   b_major = 0:              B stored as [n][k], row stride ldb
 *   b_major = 1:              B stored as [k][n], row stride ldb   (a Keras Dense kernel [in,out])
 * Rows of A/B may overlap (lda < k) — this is how strided Conv1D windows are fed without im2col.
 * Two batch dims (batch1 fastest) cover (head, batch) for attention; a batch stride of 0 broadcasts.
 * in_dtype TS_BF16 runs on tcgen05 tensor cores (TMA-fed, TMEM accumulators, fp32 accumulate);
 * in_dtype TS_F32 runs the fp32 CUDA-core engine used by the 1e-5 parity mode.
 */
typedef struct {
  const void* a;
  const void* b;
  void* c;
  int32_t m, n, k;
  int32_t a_major, b_major;
  int64_t lda, ldb, ldc;
  int32_t batch1, batch2;
  int64_t a_bs1, a_bs2, b_bs1, b_bs2, c_bs1, c_bs2;
  int32_t in_dtype;  /* ts_dtype of A and B */
  int32_t out_dtype; /* ts_dtype of C (and residual) */
  float alpha;
  const float* bias; /* [n] fp32 or NULL */
  int32_t act;       /* 0 none, 1 exact-erf GELU, 2 multiply by GELU'(act_aux[m,n]) (backward of a GELU whose input was kept) */
  const void* residual; /* same dtype/shape as C or NULL; added after act */
  int64_t ldr, r_bs1, r_bs2;
  int32_t accumulate; /* 1: C += result (only with out_dtype TS_F32, no act/residual) */
  void* c_preact;     /* optional second output: value before act (same dtype/ld as C) or NULL */
  int32_t force_engine; /* 0 auto, 1 force CUDA-core engine, 2 force tcgen05 (error if impossible), 3 force tcgen05 CTA-pair (cta_group::2) tiles, 4 force the 4-CTA cluster form (two pairs, B operand TMA-multicast) */
  float drop;           /* dropout rate applied after act (0 = off); keep-mask = hash(seed, element offset in C) */
  uint64_t seed;
  int64_t bias_bs1;     /* bias element stride per batch1 index (grouped conv: one bias slice per group) */
  const void* act_aux;  /* act == 2: the GELU input u, same dtype and batch strides as C, row stride ld_aux */
  int64_t ld_aux;
  /* GroupNormalization statistics taken in the epilogue (V:167-176: moments over time x channels-of-the-group): when gn_accum is
   * non-NULL the kernel adds, for every output row r = b * gn_rows_per_batch + t with t < gn_valid_rows, the sums of C[r, cols of
   * group g] and of their squares (fp32 accumulator values) to gn_accum[(b * gn_groups + g) * 2 + {0, 1}] (fp64 atomics; the
   * caller zeroes it before and turns it into mean / rstd after). tcgen05 engine only, bf16 or fp32 C through the TMA epilogue,
   * no act / residual / batching, n / gn_groups a multiple of 32. Saves the separate pass that re-reads the conv output. */
  double* gn_accum;
  int32_t gn_rows_per_batch, gn_valid_rows, gn_groups;
} ts_gemm_desc;

int ts_gemm(ts_ctx* ctx, const ts_gemm_desc* d, void* stream);

/* ---- K10: fused attention (tcgen05), head_dim 64, bf16 -------------------------------------------------------
 * Replaces  softmax(q k^T * scale + mask) -> dropout -> . v  and its autodiff transpose:
 *   W:147-167 (MultiHeadAttention.call: self, cross and the "anti-causal" decoder mask of W:416-418 / W:150-154),
 *   V:348-362 (Wav2Vec2MultiHeadAttention.call). Score/probability tensors never reach HBM.
 * q [B, Tq, *] / k, v [B, Tk, *] / o, d_o [B, Tq, *]: head h occupies columns [64h, 64h+64) from the given base;
 * *_ld is the row stride and *_bs the batch stride in elements (so q/k/v may be slices of one fused projection).
 * stats [B, heads, Tq, 2] fp32 (row max, log row-sum) is written by forward and read by backward;
 * dsum [B, heads, Tq] fp32 is scratch of backward. mask_mode: 0 none, 1 adds -1e9 (fp32) to keys j <= i.
 * o_lo (optional, same layout as o): forward stores the bf16 rounding residual of o there and backward adds it back when
 * forming D = rowsum(dO o O), which removes the systematic error a bf16-rounded O puts into dQ / dK.
 * Dropout keep-masks are a pure function of (seed, b, h, i, j); backward regenerates them.
 */
typedef struct {
  const void *q, *k, *v;
  void* o;
  int64_t q_ld, q_bs, kv_ld, kv_bs, o_ld, o_bs;
  float* stats;
  int32_t batch, heads, tq, tk, head_dim;
  float scale;
  int32_t mask_mode;
  float drop;
  uint64_t seed;
  /* backward only */
  const void* d_o;
  void *dq, *dk, *dv;
  int64_t dq_ld, dq_bs, dkv_ld, dkv_bs;
  float* dsum;
  void* o_lo;
  /* backward, optional: fp32 scratch [batch, tq, heads * 64]. When given, backward runs as ONE fused kernel (dQ partials are
   * reduce-added here by TMA, then rounded to dq); without it the two-kernel backward (dQ | dK, dV) is used. */
  float* dq_accum;
} ts_attn_desc;
int ts_attn_fwd(ts_ctx* ctx, const ts_attn_desc* d, void* stream);
int ts_attn_bwd(ts_ctx* ctx, const ts_attn_desc* d, void* stream);

/* ---- K1: log-mel front end ---------------------------------------------------------------------------------------
 * Replaces extract_fbank_features (W:739-766): tf.signal.stft(x, 400, 160, fft_length=400) (periodic Hann, no end padding),
 * power, tf.signal.linear_to_mel_weight_matrix(80, 201, 16000, 0, 8000), log(mel + 1e-6). One kernel; fp32 arithmetic.
 * wave [batch, n_samples] fp32 with batch stride `wave_batch_stride` (elements); frames F = ts_logmel_num_frames(n_samples).
 * out: [batch, F, 80] (mel_major = 0, the reference's return layout) or [batch, 80, F] (mel_major = 1, the layout
 * WhisperEncoder.call consumes, W:326-329); out_dtype TS_F32 or TS_BF16.
 */
int ts_logmel_num_frames(int n_samples);
int ts_logmel(ts_ctx* ctx, const float* wave, int64_t wave_batch_stride, int batch, int n_samples, void* out, int out_dtype,
              int mel_major, void* stream);

/* ---- K11 / K5: normalisation layers as single operators ---------------------------------------------------------------
 * ts_layernorm_fwd/bwd: tf.keras.layers.LayerNormalization(epsilon=1e-5) over the last axis of x [rows, cols]
 *   (W:214,216,245,249,253,322,392; V:280,411,415,554,778). dtype = TS_F32 | TS_BF16 for x / y / dy / dx; gamma, beta,
 *   mean, rstd and the parameter gradients are fp32. bwd ADDS into dgamma / dbeta; dres (optional) is added to dx.
 * ts_groupnorm_gelu_fwd: GroupNormalization(groups) + exact-erf GELU of the conv feature encoder (V:132-196, V:248-249) on
 *   x [batch, T, C]: statistics over (T, C/groups) per (batch, group); writes mean/rstd [batch, groups] and y = gelu(gn(x)).
 *   accum is scratch of 2 * batch * groups doubles. accum == NULL: mean / rstd are INPUTS (the moments were taken by the producer
 *   of x — conv0's store loop or the conv GEMM epilogue, ts_gemm_desc.gn_accum — as the train step does), x is read once.
 */
int ts_layernorm_fwd(ts_ctx* ctx, int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int rows, int cols, float eps, void* stream);
int ts_layernorm_bwd(ts_ctx* ctx, int dtype, const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int cols, void* stream);
int ts_groupnorm_gelu_fwd(ts_ctx* ctx, int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean,
                          float* rstd, double* accum, int batch, int t, int c, int groups, float eps, void* stream);
/* GroupNormalization.call alone (V:167-196), same arguments, no activation. */
int ts_groupnorm_fwd(ts_ctx* ctx, int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     double* accum, int batch, int t, int c, int groups, float eps, void* stream);

/* ---- span masking utilities ------------------------------------------------------------------------
 * apply_time_mask (V:1073-1095, axis = 1) / apply_feature_mask (V:1098-1120, axis = 2) on x [batch, t, h]:
 *   expanded[b, p] = OR_{i < mask_length} start_mask[b, p - i]   (p = time step or feature index; bit-exact integer work)
 *   y = x * (1 - expanded)
 * start_mask: uint8 [batch, t] (axis 1) or [batch, h] (axis 2) — the span starts the reference draws with
 * tf.random.uniform(...) < mask_prob; expanded_mask (fp32 0/1, same shape) is an output.
 */
int ts_span_mask_apply(ts_ctx* ctx, int dtype, const void* x, const unsigned char* start_mask, void* y, float* expanded_mask,
                       int batch, int t, int h, int axis, int mask_length, void* stream);

/* Negative sample positions of the contrastive loss, _sample_negative_indices (V:907-937): per batch row the positions of the
 * k = max(min(num_negatives, t - 1), 1) smallest of the t uniform ints `random_ints[b, :]` (smallest first, ties by lower index =
 * tf.nn.top_k(-float(r))), tiled to num_negatives entries. out int32 [batch, num_negatives] — the same list serves every time
 * step (V:933-935). Integer work, bit-exact; the random ints are an input (the caller's generator). */
int ts_w2v_sample_negatives(ts_ctx* ctx, const int32_t* random_ints, int batch, int t, int num_negatives, int32_t* out, void* stream);

/* ---- tf.nn.ctc_loss as Wav2Vec2ForCTC._compute_ctc_loss calls it (speech_jobs/whisper_single.py:897-929; SURVEY f-2) -------------
 * logits fp32 [batch, t, vocab] (batch-major: the reference's transpose to time-major is a view), labels int32 [batch, label_len]
 * dense, label_length = number of labels > 0 (WS:907), logit_length = t, `blank` = blank_index (0 there). loss_per_sample [batch]
 * = -log p(labels | logits) (+inf when no alignment exists; 0 instead with zero_infinity, WS:920-921); dlogits (nullable, dtype
 * grad_dtype, same shape as logits) = d loss_b / d logits * grad_scale — the caller folds the reduction ("sum": 1, "mean":
 * 1 / batch, WS:924-927) and the replica divisor into grad_scale. workspace: ts_ctc_workspace_floats(...) floats. */
int64_t ts_ctc_workspace_floats(int batch, int t, int label_len);
int ts_ctc_loss(ts_ctx* ctx, int grad_dtype, const float* logits, const int32_t* labels, int batch, int t, int vocab, int label_len, int blank,
                float* workspace, float* loss_per_sample, void* dlogits, float grad_scale, int zero_infinity, void* stream);

/* ---- K19/K20: gradient clipping + Keras-2.10 legacy Adam over a flat arena ------------------------
 * Replaces tf.clip_by_global_norm (V:1243, VS:1171), the optimizer's clipnorm=1.0 (V:1274, VS:1206) and
 * tf.keras.optimizers.Adam.apply_gradients (W:834, V:1246, VS:1174, WS:1179) minus its all-reduce.
 * A variable is a rows x cols block with row stride ld at `offset` (elements) inside the arenas.
 *   ts_optim_clip_global : grads *= clip / max(||grads||_2, clip) over ALL variables (local, pre-reduce).
 *   ts_optim_step        : [per-variable clipnorm] + Adam:  m += (g-m)(1-b1); v += (g*g-v)(1-b2);
 *                          p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps); optionally refreshes the bf16
 *                          compute copy of the parameters in the same pass. step >= 1 is Keras' `iterations + 1`;
 *                          step = 0 reads it from the device step state (see ts_step_state_advance).
 * fuse_global_clip = 1 folds the global-norm clip into the step (single-replica path: no all-reduce between).
 */
typedef struct ts_optim ts_optim;
int ts_optim_create(ts_ctx* ctx, int32_t n_vars, const int64_t* offsets, const int32_t* rows, const int32_t* cols,
                    const int64_t* lds, int64_t arena_elems, ts_optim** out);
void ts_optim_destroy(ts_optim* o);
int ts_optim_clip_global(ts_optim* o, float* grads, float clip, float* norm_out_dev /*nullable*/, void* stream);
/* same scale, NOT applied: scale_out_dev[0] = clip / max(||grads||_2, clip). For callers that fold the factor into the
 * collective (NCCL pre-multiplied sum) instead of spending a pass over the arena. */
int ts_optim_global_clip_scale(ts_optim* o, const float* grads, float clip, float* scale_out_dev, void* stream);
int ts_optim_step(ts_optim* o, float* params, const float* grads, float* m, float* v, void* params_bf16 /*nullable*/,
                  float lr, float beta1, float beta2, float eps, int32_t step, float global_clip /*<=0 off*/,
                  float clipnorm /*<=0 off*/, int32_t fuse_global_clip, void* stream);
/* The same step reading the gradients from a bf16 arena of the same layout — the all-reduced gradient bucket of the bf16
 * data-parallel step (ts_grad_pack_bf16 + ts_comm_allreduce_bucket): saves the unpack pass (6 B/param) and 2 B/param of gradient
 * reads; every gradient value is widened exactly: the same arithmetic as ts_grad_unpack_bf16 followed by ts_optim_step. */
int ts_optim_step_lp(ts_optim* o, float* params, const void* grads_bf16, float* m, float* v, void* params_bf16 /*nullable*/,
                     float lr, float beta1, float beta2, float eps, int32_t step, float global_clip /*<=0 off*/,
                     float clipnorm /*<=0 off*/, int32_t fuse_global_clip, void* stream);
int ts_cast_f32_to_bf16(ts_ctx* ctx, const float* src, void* dst, int64_t n, void* stream);
/* Gradient buckets for the cross-replica SUM (the all-reduce inside apply_gradients, W:834 / V:1246) in bf16: pack = bf16(src *
 * scale_dev[0]) (scale_dev nullable; the Wav2Vec2 step folds its local clip_by_global_norm factor, V:1243, in here), unpack =
 * back to the fp32 gradient arena the optimizer reads. Halves the bytes on NVLink; used when the model computes in bf16
 * (TETHYS_AR_DTYPE=fp32 keeps fp32 buckets), never in fp32 parity mode. */
int ts_grad_pack_bf16(ts_ctx* ctx, const float* src, void* dst_bf16, int64_t n, const float* scale_dev, void* stream);
int ts_grad_unpack_bf16(ts_ctx* ctx, const void* src_bf16, float* dst, int64_t n, void* stream);
/* ---- K21-K23: the collective layer (NCCL over NVLink 5 / NVSwitch, called directly) --------------------------------------
 * Replaces what tf.distribute.MultiWorkerMirroredStrategy does for the step: the gradient all-reduce inside
 * optimizer.apply_gradients (W:834, V:1246), strategy.reduce(SUM) of the losses (W:848, V:1260), the broadcast of the chief's
 * weights under strategy.scope() (W:896, V:1266), CommunicationOptions(NCCL, timeout 120 s) (V:1463-1475).
 *   ts_comm_unique_id   rank 0 draws the 128-byte NCCL id; the HOST hands it to the other ranks (store / file / MPI).
 *   ts_comm_init        ncclCommInitRank on the context's device. One communicator per (process, GPU).
 *   ts_comm_alloc/free  ncclMemAlloc memory registered with the communicator (ncclCommRegister): buffers the switch can reduce
 *                       in place (NVLS); the gradient arenas live there. Falls back to cudaMalloc on an NCCL without it.
 *   ts_comm_broadcast   in place, from `root`.
 *   ts_comm_allreduce_bucket  in-place SUM of `count` elements (TS_F32 / TS_BF16). premul_scale_dev (nullable): a device scalar
 *                       of the same dtype; this rank contributes premul * x (ncclRedOpCreatePreMulSum) — the local
 *                       clip_by_global_norm factor of V:1243 at no extra pass.
 *   ts_comm_check       ncclCommGetAsyncError -> TS_ENCCL with the NCCL message (a hung / failed collective becomes an error).
 *   ts_comm_finalize    frees the registered buffers and destroys the communicator.
 * All collectives are enqueued on the caller's stream (stream-ordered, capturable into the step's CUDA graph). libnccl is
 * resolved with dlopen at the first call (TETHYS_NCCL_LIB overrides the search). */
typedef struct ts_comm ts_comm;
int ts_comm_unique_id(ts_ctx* ctx, void* out128);
int ts_comm_init(ts_ctx* ctx, const void* unique_id128, int nranks, int rank, ts_comm** out);
int ts_comm_info(ts_comm* c, int* nranks, int* rank, int* nccl_version, int* registered_buffers);
int ts_comm_alloc(ts_comm* c, int64_t bytes, void** ptr);
int ts_comm_free(ts_comm* c, void* ptr);
int ts_comm_broadcast(ts_comm* c, void* buf, int64_t count, int dtype, int root, void* stream);
int ts_comm_allreduce_bucket(ts_comm* c, void* buf, int64_t count, int dtype, const void* premul_scale_dev, void* stream);
int ts_comm_check(ts_comm* c);
int ts_comm_finalize(ts_comm* c);

/* ---- composite train step ("whole fwd + bwd + reduce + update on pre-bound arenas") -------------------------
 * One call = one step of the reference's train loop body on the arenas given to ts_*_bind:
 *   ts_w2v_step      train_step VS:1119-1176 (comm == NULL) / distributed_train_step V:1186-1260 (comm != NULL): forward with
 *                    loss / N (V:1231), backward, LOCAL clip_by_global_norm (V:1243, VS:1171), all-reduce SUM (inside
 *                    apply_gradients, V:1246), per-variable clipnorm (V:1274) + Adam, return value sum_r loss_r / N (V:1260).
 *   ts_whisper_step  distributed_train_step W:819-848: forward + shifted CE, backward, all-reduce SUM NOT divided by N (W:834),
 *                    Adam (W:901: no clipping -> global_clip = clipnorm = 0), return value sum_r loss_r (W:848).
 * Everything, the NCCL collectives included, is enqueued on `stream`: the call is asynchronous and may be captured in a CUDA
 * graph (use step = 0 and ts_step_state_advance then). The caller keeps the bf16 compute weights in sync before the first
 * step (ts_*_sync_compute_weights); every step refreshes them in its update pass. Activations, logits and the per-replica
 * loss stay readable through ts_*_get_buffer afterwards. */
typedef struct ts_step_args {
  ts_optim* optim;        /* over the program's variable table (ts_*_param_info), see ts_optim_create */
  float* adam_m;          /* fp32 [arena_elems] */
  float* adam_v;          /* fp32 [arena_elems] */
  float lr, beta1, beta2, eps;
  int32_t step;           /* Keras `iterations + 1` (>= 1); 0 = read it from the device step state */
  float global_clip;      /* tf.clip_by_global_norm on this replica's gradients, before the reduce; <= 0: none */
  float clipnorm;         /* the optimizer's per-variable clipnorm, after the reduce; <= 0: none */
  uint64_t seed;          /* dropout seed of this step */
  int32_t dropout;        /* 0: dropout layers off (parity runs) */
  int32_t reserved;
  ts_comm* comm;          /* NULL: one replica */
  void* grads_bf16;       /* with comm, optional: bf16 [arena_elems] (ts_comm_alloc) - the gradients cross the link as a bf16
                             bucket and Adam reads the reduced bucket directly; NULL: the fp32 gradient arena is reduced in place */
  float* scratch_dev;     /* with comm and global_clip > 0: 2 floats of device scratch (the local clip factor) */
  float* loss_out_dev;    /* nullable: device scalar that receives the step's return value */
} ts_step_args;

/* tf.keras.layers.Dropout in training mode (W:160, W:203-205, W:342; V:281, V:393-396, V:431) over a flat tensor:
 * y[i] = x[i] * mask(seed, i) / (1 - rate), in place allowed. mask is the library's counter-based generator (element i of the
 * tensor -> chunk i >> 5, position i & 31), the same one the GEMM epilogues (ts_gemm_desc.drop / .seed) and every backward
 * kernel evaluate, so a mask can be regenerated anywhere from (seed, flat index) and is never stored. */
int ts_dropout(ts_ctx* ctx, int dtype, const void* x, void* y, int64_t n, float rate, uint64_t seed, void* stream);

/* ---- Wav2Vec2 pre-training program ---------------------------------------------------------------------
 * Replaces Wav2Vec2ForPreTraining.call(training=True) + _compute_contrastive_loss + _compute_diversity_loss
 * + tape.gradient of the per-replica step (V:768-905, V:1199-1240; VS:1131-1168; WS:1143-1176).
 * The negative indices are an input (the TF RNG stream of V:919 cannot be reproduced): neg[b*neg_bs + t*neg_ts + k].
 * Parameters/gradients live in caller-owned flat fp32 arenas whose layout ts_w2v_param_info describes; the
 * arena is ordered so that after backward stage s all gradients below ts_w2v_stage_end(s) are final
 * (stage 0 = heads, 1..L = encoder layers L-1..0, L+1 = feature projection + conv front end), which is what
 * the host uses to overlap the bucketed gradient all-reduce (K21) with backward.
 */
typedef struct {
  int32_t hidden, layers, heads, ffn;
  int32_t n_conv;
  int32_t conv_dim[8], conv_kernel[8], conv_stride[8];
  int32_t pos_kernel, pos_groups; /* num_conv_pos_embeddings / _groups; pos_groups is also the GroupNorm group count (V:248) */
  int32_t cv_groups, cv_per_group, cv_dim, proj_dim;
  int32_t num_negatives;
  float ln_eps, temperature, diversity_weight;
  float hidden_dropout, activation_dropout, attention_dropout;
  /* task head on the trunk (SURVEY f-2): 0 = pre-training (contrastive), 1 = Wav2Vec2ForCTC (V:940-1001),
   * 2 = Wav2Vec2ForSequenceClassification (V:1004-1070). Heads 1/2 have no project_hid / project_q variables (Keras never
   * builds them there); the quantizer's exist and get zero gradients, as in the reference (VS:1163-1166). */
  int32_t head;
  int32_t vocab_size;       /* head 1: lm_head width (V:104) */
  int32_t classifier_proj;  /* head 2: classifier_proj_size (V:108-114) */
  int32_t num_labels;       /* head 2: VS:131 */
} ts_w2v_config;

typedef struct ts_w2v ts_w2v;
int ts_w2v_create(ts_ctx* ctx, const ts_w2v_config* cfg, int precision /*TS_F32 | TS_BF16*/, ts_w2v** out);
void ts_w2v_destroy(ts_w2v* m);
int64_t ts_w2v_arena_elems(ts_w2v* m);
int ts_w2v_num_params(ts_w2v* m);
int ts_w2v_param_info(ts_w2v* m, int i, char* name, int name_cap, int64_t* offset, int32_t* ndim, int64_t* shape4,
                      int64_t* ld);
int ts_w2v_num_stages(ts_w2v* m);
int64_t ts_w2v_stage_end(ts_w2v* m, int stage);
int64_t ts_w2v_workspace_bytes(ts_w2v* m, int batch, int n_samples);
int ts_w2v_bind(ts_w2v* m, float* params, float* grads, void* params_bf16 /*bf16 mode*/, void* workspace,
                int64_t workspace_bytes);
int ts_w2v_sync_compute_weights(ts_w2v* m, void* stream); /* fp32 master -> bf16 compute copy */
int ts_w2v_forward(ts_w2v* m, const float* wave /*[B,N]*/, int batch, int n_samples, const int32_t* neg, int64_t neg_bs,
                   int64_t neg_ts, float loss_div /*num replicas, V:1231*/, uint64_t seed, int training, void* stream);
/* heads 1/2: model(features, labels=labels, training=…) of VS:1156-1157. labels: int32 [batch] for head 2 (NULL = zeros, the
 * dummy dataset's label, V:1139), ignored by head 1 (its stand-in "CTC" loss targets class 0 on every frame, V:997-1000).
 * training != 0 also runs the quantizer (V:784-789) and computes loss ("scalars"[0]) + the loss gradient for
 * ts_w2v_backward; dropout != 0 enables the dropout layers of a training call (parity runs switch them off); logits are in
 * the "head_logits" buffer (fp32). */
int ts_w2v_forward_head(ts_w2v* m, const float* wave, int batch, int n_samples, const int32_t* labels, float loss_div,
                        uint64_t seed, int dropout, int training, void* stream);
/* Wav2Vec2FeatureExtractor.call only (V:283-298): conv stack + GroupNorm/GELU + positional conv + LayerNorm, inference mode;
 * result in the "extract_features" buffer. Used by the front-end microbench (BASELINE config 5). */
int ts_w2v_forward_features(ts_w2v* m, const float* wave, int batch, int n_samples, void* stream);
int ts_w2v_backward(ts_w2v* m, int stage_from, int stage_to, void* stream);
/* named views into the workspace of the last forward ("scalars" = {loss, contrastive, perplexity, raw sum}). */
int ts_w2v_get_buffer(ts_w2v* m, const char* name, void** ptr, int32_t* dtype, int32_t* ndim, int64_t* shape4);

/* Pre-training program (cfg.head == 0): neg / neg_bs / neg_ts as in ts_w2v_forward, labels ignored. Task-head programs
 * (cfg.head 1 / 2, VS:1155-1157): labels as in ts_w2v_forward_head, neg ignored. */
int ts_w2v_step(ts_w2v* m, const float* wave, int batch, int n_samples, const int32_t* neg, int64_t neg_bs, int64_t neg_ts,
                const int32_t* labels, const ts_step_args* args, void* stream);

/* ---- Whisper encoder-decoder program ---------------------------------------------------------------------
 * Replaces WhisperForConditionalGeneration.call(features, labels=labels, training=True) and tape.gradient of
 * distributed_train_step's per-replica body (W:547-616, W:826-833). features [B, n_mels, T_mel] fp32,
 * labels [B, S] int32. Bug-compatible with the reference (anti-causal decoder mask, double label shift).
 * Backward stages: 0 lm_head + final decoder LN | 1..Ld decoder layers Ld-1..0 | Ld+1 embedding + final encoder LN |
 * Ld+2..Ld+1+Le encoder layers Le-1..0 | Ld+Le+2 conv stem.
 */
typedef struct {
  int32_t d_model, enc_layers, dec_layers, heads, d_ff, n_mels, n_ctx, vocab, max_target, start_token;
  float ln_eps, dropout, attention_dropout, activation_dropout;
} ts_whisper_config;

typedef struct ts_whisper ts_whisper;
int ts_whisper_create(ts_ctx* ctx, const ts_whisper_config* cfg, int precision, ts_whisper** out);
void ts_whisper_destroy(ts_whisper* m);
int64_t ts_whisper_arena_elems(ts_whisper* m);
int ts_whisper_num_params(ts_whisper* m);
int ts_whisper_param_info(ts_whisper* m, int i, char* name, int name_cap, int64_t* offset, int32_t* ndim, int64_t* shape4,
                          int64_t* ld);
int ts_whisper_num_stages(ts_whisper* m);
int64_t ts_whisper_stage_end(ts_whisper* m, int stage);
int64_t ts_whisper_workspace_bytes(ts_whisper* m, int batch, int t_mel, int seq);
int ts_whisper_bind(ts_whisper* m, float* params, float* grads, void* params_bf16, void* workspace, int64_t workspace_bytes);
int ts_whisper_sync_compute_weights(ts_whisper* m, void* stream);
int ts_whisper_forward(ts_whisper* m, const float* features, int batch, int t_mel, const int32_t* labels, int seq, uint64_t seed,
                       int training, int compute_loss, void* stream);
/* Greedy decoding — WhisperForConditionalGeneration.generate (W:636-709). ts_whisper_encode runs the encoder once in
 * inference mode and projects the cross-attention keys/values of every decoder layer (the only part of the decoder state
 * that can be cached: under the reference's anti-causal self-attention mask, W:414-418, earlier positions attend to later
 * ones, so every step re-runs the decoder over the whole prefix, exactly as generate() does). ts_whisper_decode_step runs
 * the decoder on ids = [decoder_start_token, tokens[b, 0 .. len-2]] and writes tokens[b, len-1] = argmax over the vocabulary
 * of the last position's logits (first maximum, tf.argmax); 1 <= len <= max_len; `tokens` is int32 [batch, ld_tok] on the
 * device. The last step's logits are the "next_token_logits" buffer. */
int ts_whisper_encode(ts_whisper* m, const float* feats, int batch, int n_frames, int max_len, void* stream);
int ts_whisper_decode_step(ts_whisper* m, int32_t* tokens, int64_t ld_tok, int len, void* stream);
int ts_whisper_backward(ts_whisper* m, int stage_from, int stage_to, void* stream);
int ts_whisper_get_buffer(ts_whisper* m, const char* name, void** ptr, int32_t* dtype, int32_t* ndim, int64_t* shape4);
int ts_whisper_step(ts_whisper* m, const float* features, int batch, int t_mel, const int32_t* labels, int seq,
                    const ts_step_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TETHYS_H_ */
