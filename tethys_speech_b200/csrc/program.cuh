// Shared plumbing of the model programs (Wav2Vec2 / Whisper train steps): parameter table over a flat arena,
// workspace bump allocator, GEMM descriptor builder.
#pragma once
#include <string>
#include <vector>
#include <string.h>
#include "ops.cuh"

namespace ts {

struct ParamDef {
  std::string name;
  long long offset;   // element offset in the fp32 arenas (params / grads / m / v) and the bf16 compute copy
  int ndim;
  long long shape[4];
  int rows, cols;     // 2D segment view: rows x cols with row stride ld (dense: rows = 1, cols = numel)
  long long ld;
};

struct ParamTable {
  std::vector<ParamDef> defs;
  long long n = 0;  // arena elements
  // dense tensor, 64-element aligned
  long long add(const std::string& name, std::initializer_list<long long> shape) {
    ParamDef d;
    d.name = name;
    d.ndim = (int)shape.size();
    long long numel = 1;
    int i = 0;
    for (long long s : shape) { d.shape[i++] = s; numel *= s; }
    for (; i < 4; ++i) d.shape[i] = 1;
    n = (n + 63) & ~63ll;
    d.offset = n;
    d.rows = 1;
    d.cols = (int)numel;
    d.ld = numel;
    n += numel;
    defs.push_back(d);
    return d.offset;
  }
  // `parts` matrices [rows, cols_each] fused side by side into one [rows, parts*cols_each (+pad)] block
  long long add_fused(const std::vector<std::string>& names, long long rows, long long cols_each, long long ld) {
    n = (n + 63) & ~63ll;
    const long long base = n;
    for (size_t j = 0; j < names.size(); ++j) {
      ParamDef d;
      d.name = names[j];
      d.ndim = rows > 1 ? 2 : 1;
      d.shape[0] = rows > 1 ? rows : cols_each;
      d.shape[1] = rows > 1 ? cols_each : 1;
      d.shape[2] = d.shape[3] = 1;
      d.offset = base + (long long)j * cols_each;
      d.rows = (int)rows;
      d.cols = (int)cols_each;
      d.ld = ld;
      defs.push_back(d);
    }
    n += rows * ld;
    return base;
  }
};

struct Bump {
  char* base = nullptr;
  size_t off = 0;
  void* get(size_t bytes) {
    const size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    return base ? base + a : nullptr;
  }
};

struct GemmB {
  ts_gemm_desc d;
  GemmB(int in_dt, int out_dt) {
    memset(&d, 0, sizeof(d));
    d.in_dtype = in_dt; d.out_dtype = out_dt; d.alpha = 1.f; d.batch1 = 1; d.batch2 = 1;
  }
  GemmB& A(const void* p, int major, long long ld) { d.a = p; d.a_major = major; d.lda = ld; return *this; }
  GemmB& B(const void* p, int major, long long ld) { d.b = p; d.b_major = major; d.ldb = ld; return *this; }
  GemmB& C(void* p, long long ld) { d.c = p; d.ldc = ld; return *this; }
  GemmB& mnk(int m, int n, int k) { d.m = m; d.n = n; d.k = k; return *this; }
  GemmB& batch(int b1, int b2) { d.batch1 = b1; d.batch2 = b2; return *this; }
  GemmB& astride(long long s1, long long s2) { d.a_bs1 = s1; d.a_bs2 = s2; return *this; }
  GemmB& bstride(long long s1, long long s2) { d.b_bs1 = s1; d.b_bs2 = s2; return *this; }
  GemmB& cstride(long long s1, long long s2) { d.c_bs1 = s1; d.c_bs2 = s2; return *this; }
  GemmB& bias(const float* b, long long bs1 = 0) { d.bias = b; d.bias_bs1 = bs1; return *this; }
  GemmB& gelu(void* preact) { d.act = 1; d.c_preact = preact; return *this; }
  GemmB& gelu_grad(const void* u, long long ld) { d.act = 2; d.act_aux = u; d.ld_aux = ld; return *this; }
  GemmB& res(const void* r, long long ld, long long s1 = 0, long long s2 = 0) { d.residual = r; d.ldr = ld; d.r_bs1 = s1; d.r_bs2 = s2; return *this; }
  GemmB& alpha(float a) { d.alpha = a; return *this; }
  GemmB& drop(float rate, uint64_t seed) { d.drop = rate; d.seed = seed; return *this; }
  GemmB& acc() { d.accumulate = 1; return *this; }
  GemmB& gn_stats(double* accum, int rows_per_batch, int valid_rows, int groups) {
    d.gn_accum = accum; d.gn_rows_per_batch = rows_per_batch; d.gn_valid_rows = valid_rows; d.gn_groups = groups; return *this;
  }
  GemmB& simt() { d.force_engine = 1; return *this; }   // tiny / oddly strided problems: CUDA-core engine
  int run(Ctx* ctx, cudaStream_t st) { return gemm(ctx, &d, st); }
};

#define TS_TRY(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

static inline uint64_t site_seed(uint64_t base, uint64_t site) { return base * 0x9E3779B97F4A7C15ull + site * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull; }

static inline void same_pad(int t_in, int k, int s, int* t_out, int* left, int* right) {
  const int to = (t_in + s - 1) / s;
  int pt = (to - 1) * s + k - t_in;
  if (pt < 0) pt = 0;
  *t_out = to; *left = pt / 2; *right = pt - pt / 2;
}

// ---- composite train step (ts_w2v_step / ts_whisper_step) ---------------------------------------------------
// Everything after backward: [local clip_by_global_norm factor] -> cross-replica SUM of the gradient arena -> per-variable
// clipnorm + Keras-legacy Adam (+ refresh of the bf16 compute weights) -> the step's return value. All of it is enqueued on
// `st` (NCCL included), so a caller may capture the whole step in one CUDA graph.
//   loss_mean_over_replicas: Wav2Vec2's convention (V:1231, V:1260: sum_r loss_r / N); Whisper returns sum_r loss_r (W:848).
static inline int step_reduce_update(Ctx* ctx, float* P, float* G, void* P16, long long n, const float* loss_dev,
                                     bool loss_mean_over_replicas, const ts_step_args* a, cudaStream_t st) {
  TS_REQUIRE(ctx, a && a->optim && a->adam_m && a->adam_v, TS_EINVAL, "step: optimizer handle and Adam state arenas are required");
  int nranks = 1;
  if (a->comm) {
    int rank = 0, ver = 0, reg = 0;
    TS_TRY(ts_comm_info(a->comm, &nranks, &rank, &ver, &reg));
  }
  if (!a->comm) {
    // one replica: clip_by_global_norm folded into the update pass (VS:1171-1174)
    TS_TRY(ts_optim_step(a->optim, P, G, a->adam_m, a->adam_v, P16, a->lr, a->beta1, a->beta2, a->eps, a->step, a->global_clip,
                         a->clipnorm, a->global_clip > 0.f ? 1 : 0, st));
  } else {
    const float* scale = nullptr;
    if (a->global_clip > 0.f) {   // V:1243: the clip is LOCAL and happens before the reduce
      TS_REQUIRE(ctx, a->scratch_dev, TS_EINVAL, "step: scratch_dev (2 floats) is required with comm and global_clip");
      TS_TRY(ts_optim_global_clip_scale(a->optim, G, a->global_clip, a->scratch_dev, st));
      scale = a->scratch_dev;
    }
    if (a->grads_bf16) {          // bf16 bucket: the clip factor rides on the pack; Adam reads the reduced bucket as it is
      TS_TRY(grad_pack_bf16(ctx, G, a->grads_bf16, n, scale, st));
      TS_TRY(ts_comm_allreduce_bucket(a->comm, a->grads_bf16, n, TS_BF16, nullptr, st));
      TS_TRY(ts_optim_step_lp(a->optim, P, a->grads_bf16, a->adam_m, a->adam_v, P16, a->lr, a->beta1, a->beta2, a->eps, a->step, 0.f,
                              a->clipnorm, 0, st));
    } else {                      // fp32 arena in place; sum_r scale_r * g_r as one pre-multiplied sum
      TS_TRY(ts_comm_allreduce_bucket(a->comm, G, n, TS_F32, scale, st));
      TS_TRY(ts_optim_step(a->optim, P, G, a->adam_m, a->adam_v, P16, a->lr, a->beta1, a->beta2, a->eps, a->step, 0.f, a->clipnorm, 0, st));
    }
  }
  if (a->loss_out_dev) {
    TS_CUDA_OK(ctx, cudaMemcpyAsync(a->loss_out_dev, loss_dev, sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (a->comm) {
      if (loss_mean_over_replicas && nranks > 1) TS_TRY(scale_inplace(ctx, a->loss_out_dev, 1, nullptr, 1.f / (float)nranks, st));
      TS_TRY(ts_comm_allreduce_bucket(a->comm, a->loss_out_dev, 1, TS_F32, nullptr, st));
    }
  }
  return 0;
}

}  // namespace ts
