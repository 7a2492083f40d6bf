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

/* ---- K9: GEMM with fused epilogue ------------------------------------------------------------
 * Replaces every tf.keras.layers.Dense / tf.matmul / Conv1D-as-GEMM on the path:
 *   W:89-92,141,174,194-205,311-312,545 ; V:240-268,316-319,338-340,371,383-398,553,579,586 and
 *   their autodiff transposes (dgrad, wgrad).
 *
 *   C[m,n] = act(alpha * sum_k A[m,k]*B[n,k] + bias[n]) + residual[m,n]        (per batch)
 *
 * Storage ("major") of the two operands, in elements:
 *   a_major = 0 ("K-major"):  A stored as [m][k], row stride lda   (reduce dim contiguous)
 *   a_major = 1 ("MN-major"): A stored as [k][m], row stride lda   (m contiguous)
 *   b_major = 0:              B stored as [n][k], row stride ldb
 *   b_major = 1:              B stored as [k][n], row stride ldb   (a Keras Dense kernel [in,out])
 * Rows of A/B may overlap (lda < k) — this is how strided Conv1D windows are fed without im2col.
 * Two batch dims (batch1 fastest) cover (head, batch) for attention.
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
  int32_t act;       /* 0 none, 1 exact-erf GELU */
  const void* residual; /* same dtype/shape as C or NULL; added after act */
  int64_t ldr, r_bs1, r_bs2;
  int32_t accumulate; /* 1: C += result (only with out_dtype TS_F32, no act/residual) */
  void* c_preact;     /* optional second output: value before act (same dtype/ld as C) or NULL */
  int32_t force_engine; /* 0 auto, 1 force CUDA-core engine, 2 force tcgen05 (error if impossible) */
} ts_gemm_desc;

int ts_gemm(ts_ctx* ctx, const ts_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TETHYS_H_ */
