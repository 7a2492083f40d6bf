// C-ABI: context lifecycle, error reporting, GEMM dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"
#include "ops.cuh"

namespace ts {

struct CtxHolder { Ctx c; };

bool pdl_enabled() {
  static const bool on = !(getenv("TETHYS_PDL") && atoi(getenv("TETHYS_PDL")) == 0);
  return on;
}

int set_err(Ctx* c, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

int gemm_tc(Ctx* ctx, const ts_gemm_desc* d, cudaStream_t st);
int gemm_simt(Ctx* ctx, const ts_gemm_desc* d, cudaStream_t st);
bool gemm_tc_supported(const ts_gemm_desc* d);
void tmap_cache_free(Ctx* ctx);

int gemm(Ctx* ctx, const ts_gemm_desc* d, cudaStream_t st) {
  if (d->gn_accum && (d->force_engine == 1 || d->in_dtype != TS_BF16 || !gemm_tc_supported(d)))
    return set_err(ctx, TS_EUNSUPPORTED, "gemm: epilogue GroupNorm statistics (gn_accum) exist on the tcgen05 engine only");
  if (d->force_engine == 1) return gemm_simt(ctx, d, st);
  if (d->force_engine >= 2 && d->force_engine <= 4) return gemm_tc(ctx, d, st);
  if (d->in_dtype == TS_BF16 && gemm_tc_supported(d)) return gemm_tc(ctx, d, st);
  if (d->in_dtype == TS_BF16) {
    // a bf16 GEMM that TMA cannot describe (unaligned leading dimension / pointer) runs on the CUDA-core engine: correct but
    // ~20x slower. Never silent: counted (ts_simt_downgrades) and, under TETHYS_STRICT_TC=1, an error.
    ctx->simt_downgrades++;
    static const bool strict = getenv("TETHYS_STRICT_TC") && atoi(getenv("TETHYS_STRICT_TC")) != 0;
    if (strict)
      return set_err(ctx, TS_EUNSUPPORTED, "TETHYS_STRICT_TC: bf16 GEMM m=%lld n=%lld k=%lld (lda %lld ldb %lld ldc %lld) is not "
                     "expressible as TMA tiles and would run on the CUDA-core engine", (long long)d->m, (long long)d->n,
                     (long long)d->k, (long long)d->lda, (long long)d->ldb, (long long)d->ldc);
  }
  return gemm_simt(ctx, d, st);
}

}  // namespace ts

using ts::Ctx;

extern "C" {

int ts_version(void) { return TS_VERSION; }

int ts_create(int device, ts_ctx** out) {
  if (!out) return TS_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return TS_ECUDA;
  if (cudaSetDevice(device) != cudaSuccess) return TS_ECUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TS_ECUDA;
  if (prop.major != 10) return TS_EUNSUPPORTED;  // sm_100a only: no other code path exists
  Ctx* c = new Ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  // TETHYS_SM_MARGIN=n: size persistent grids for n fewer SMs (leaves room for NCCL's CTAs when the gradient all-reduce overlaps backward)
  if (const char* mg = getenv("TETHYS_SM_MARGIN")) { const int n = atoi(mg); if (n > 0 && n < c->num_sms) c->num_sms -= n; }
  if (cudaMalloc(&c->d_watchdog, sizeof(int)) != cudaSuccess) { delete c; return TS_ECUDA; }
  cudaMemset(c->d_watchdog, 0, sizeof(int));
  if (cudaMalloc(&c->d_state, 2 * sizeof(unsigned long long)) != cudaSuccess) { cudaFree(c->d_watchdog); delete c; return TS_ECUDA; }
  cudaMemset(c->d_state, 0, 2 * sizeof(unsigned long long));
  *out = reinterpret_cast<ts_ctx*>(c);
  return TS_OK;
}

void ts_destroy(ts_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return;
  ts::tmap_cache_free(c);
  if (c->d_watchdog) cudaFree(c->d_watchdog);
  if (c->d_state) cudaFree(c->d_state);
  delete c;
}

const char* ts_last_error(ts_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  return c ? c->err.c_str() : "null context";
}

int ts_watchdog_check(ts_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return TS_EINVAL;
  int v = 0;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return ts::set_err(c, TS_ECUDA, "device error: %s", cudaGetErrorString(e));
  e = cudaMemcpy(&v, c->d_watchdog, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return ts::set_err(c, TS_ECUDA, "watchdog read failed: %s", cudaGetErrorString(e));
  if (v != 0) {
    cudaMemset(c->d_watchdog, 0, sizeof(int));
    return ts::set_err(c, TS_EWATCHDOG, "device mbarrier wait timed out (role code %d)", v);
  }
  return TS_OK;
}

int64_t ts_launch_count(ts_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  return c ? (int64_t)c->launches : 0;
}

int ts_debug_gemm_trace(ts_ctx* ctx, void* device_buf) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return TS_EINVAL;
  c->gemm_trace = device_buf;
  return 0;
}

int64_t ts_simt_downgrades(ts_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  return c ? (int64_t)c->simt_downgrades : 0;
}

__global__ void step_state_set_kernel(unsigned long long* st, unsigned long long salt, unsigned long long step) {
  ts::pdl_enter(); st[0] = salt; st[1] = step; }
__global__ void step_state_advance_kernel(unsigned long long* st) {
  ts::pdl_enter(); st[0] += 1ull; st[1] += 1ull; }

int ts_step_state_set(ts_ctx* ctx, uint64_t salt, int64_t step, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || step < 0) return TS_EINVAL;
  ts::launch_k(step_state_set_kernel, 1, 1, 0, reinterpret_cast<cudaStream_t>(stream), c->d_state, salt, (unsigned long long)step);
  TS_LAUNCH_OK(c);
  return 0;
}
int ts_step_state_get(ts_ctx* ctx, uint64_t* salt, int64_t* step) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !salt || !step) return TS_EINVAL;
  unsigned long long h[2] = {0, 0};
  TS_CUDA_OK(c, cudaDeviceSynchronize());
  TS_CUDA_OK(c, cudaMemcpy(h, c->d_state, sizeof(h), cudaMemcpyDeviceToHost));
  *salt = h[0]; *step = (int64_t)h[1];
  return 0;
}
int ts_step_state_advance(ts_ctx* ctx, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return TS_EINVAL;
  ts::launch_k(step_state_advance_kernel, 1, 1, 0, reinterpret_cast<cudaStream_t>(stream), c->d_state);
  TS_LAUNCH_OK(c);
  return 0;
}

int ts_layernorm_fwd(ts_ctx* ctx, int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int rows, int cols, float eps, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return TS_EINVAL;
  return ts::layernorm_fwd(c, dtype, x, nullptr, gamma, beta, y, nullptr, mean, rstd, rows, cols, eps, reinterpret_cast<cudaStream_t>(stream));
}
int ts_layernorm_bwd(ts_ctx* ctx, int dtype, const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int cols, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return TS_EINVAL;
  return ts::layernorm_bwd(c, dtype, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, cols, reinterpret_cast<cudaStream_t>(stream));
}
int ts_groupnorm_gelu_fwd(ts_ctx* ctx, int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean,
                          float* rstd, double* accum, int batch, int t, int ch, int groups, float eps, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return TS_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (accum) {
    int rc = ts::groupnorm_stats(c, dtype, x, accum, mean, rstd, batch, t, ch, groups, t, eps, st);
    if (rc) return rc;
  }
  return ts::groupnorm_gelu_fwd(c, dtype, x, t, mean, rstd, gamma, beta, y, t, 0, batch, t, ch, groups, st);
}

int ts_groupnorm_fwd(ts_ctx* ctx, int dtype, const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     double* accum, int batch, int t, int ch, int groups, float eps, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return TS_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = ts::groupnorm_stats(c, dtype, x, accum, mean, rstd, batch, t, ch, groups, t, eps, st);
  if (rc) return rc;
  return ts::groupnorm_fwd(c, dtype, x, mean, rstd, gamma, beta, y, batch, t, ch, groups, st);
}

int ts_gemm(ts_ctx* ctx, const ts_gemm_desc* d, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !d) return TS_EINVAL;
  return ts::gemm(c, d, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ---- optimizer object ------------------------------------------------------------------------------------
#include <vector>
#include "ops.cuh"

namespace ts {
struct Optim {
  Ctx* ctx;
  WorkItem* d_items = nullptr;   // built once here, freed in ts_optim_destroy (no global cache)
  int nitems = 0;
  int nseg = 0;
  long long arena = 0;
  float* d_sumsq = nullptr;   // [nseg]
  float* d_scal = nullptr;    // [0] global clip scale, [1] global norm
};
}  // namespace ts

extern "C" {

int ts_optim_create(ts_ctx* ctx_, int32_t n, const int64_t* offsets, const int32_t* rows, const int32_t* cols,
                    const int64_t* lds, int64_t arena_elems, ts_optim** out) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !out || n <= 0) return TS_EINVAL;
  std::vector<ts::Segment> h(n);
  for (int i = 0; i < n; ++i) {
    h[i].offset = offsets[i]; h[i].rows = rows[i]; h[i].cols = cols[i]; h[i].ld = lds[i];
    TS_REQUIRE(ctx, offsets[i] >= 0 && rows[i] > 0 && cols[i] > 0 &&
                        offsets[i] + (long long)(rows[i] - 1) * lds[i] + cols[i] <= arena_elems,
               TS_EINVAL, "optim: variable %d outside the arena", i);
  }
  ts::Optim* o = new ts::Optim();
  o->ctx = ctx; o->nseg = n; o->arena = arena_elems;
  const std::vector<ts::WorkItem> items = ts::build_work_items(h);
  o->nitems = (int)items.size();
  cudaError_t e = cudaMalloc(&o->d_items, sizeof(ts::WorkItem) * items.size());
  if (e == cudaSuccess) e = cudaMemcpy(o->d_items, items.data(), sizeof(ts::WorkItem) * items.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&o->d_sumsq, sizeof(float) * n);
  if (e == cudaSuccess) e = cudaMalloc(&o->d_scal, sizeof(float) * 4);
  if (e != cudaSuccess) {
    cudaFree(o->d_items); cudaFree(o->d_sumsq); cudaFree(o->d_scal);
    delete o;
    return ts::set_err(ctx, TS_ECUDA, "optim: device allocation failed: %s", cudaGetErrorString(e));
  }
  *out = reinterpret_cast<ts_optim*>(o);
  return 0;
}

void ts_optim_destroy(ts_optim* o_) {
  ts::Optim* o = reinterpret_cast<ts::Optim*>(o_);
  if (!o) return;
  cudaFree(o->d_items); cudaFree(o->d_sumsq); cudaFree(o->d_scal);
  delete o;
}

int ts_optim_clip_global(ts_optim* o_, float* grads, float clip, float* norm_out_dev, void* stream) {
  ts::Optim* o = reinterpret_cast<ts::Optim*>(o_);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ts::grad_sumsq(o->ctx, grads, TS_F32, o->d_items, o->nitems, o->nseg, o->d_sumsq, st);
  if (rc) return rc;
  rc = ts::global_clip_scale(o->ctx, o->d_sumsq, o->nseg, clip, o->d_scal, norm_out_dev ? norm_out_dev : o->d_scal + 1, st);
  if (rc) return rc;
  return ts::scale_inplace(o->ctx, grads, o->arena, o->d_scal, 1.f, st);
}

int ts_optim_global_clip_scale(ts_optim* o_, const float* grads, float clip, float* scale_out_dev, void* stream) {
  ts::Optim* o = reinterpret_cast<ts::Optim*>(o_);
  cudaStream_t st = (cudaStream_t)stream;
  if (!o || !grads || !scale_out_dev) return TS_EINVAL;
  int rc = ts::grad_sumsq(o->ctx, grads, TS_F32, o->d_items, o->nitems, o->nseg, o->d_sumsq, st);
  if (rc) return rc;
  return ts::global_clip_scale(o->ctx, o->d_sumsq, o->nseg, clip, scale_out_dev, o->d_scal + 1, st);
}

static int optim_step_impl(ts_optim* o_, float* params, const void* grads, int grad_dt, float* m, float* v, void* params_bf16, float lr,
                           float beta1, float beta2, float eps, int32_t step, float global_clip, float clipnorm,
                           int32_t fuse_global_clip, void* stream) {
  ts::Optim* o = reinterpret_cast<ts::Optim*>(o_);
  if (!o || !params || !grads || !m || !v) return TS_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  ts::AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.step = step; a.clipnorm = clipnorm;
  a.pre_scale = nullptr; a.sumsq = nullptr;
  if (clipnorm > 0.f || (fuse_global_clip && global_clip > 0.f)) {
    int rc = ts::grad_sumsq(o->ctx, grads, grad_dt, o->d_items, o->nitems, o->nseg, o->d_sumsq, st);
    if (rc) return rc;
    a.sumsq = o->d_sumsq;
  }
  if (fuse_global_clip && global_clip > 0.f) {
    int rc = ts::global_clip_scale(o->ctx, o->d_sumsq, o->nseg, global_clip, o->d_scal, o->d_scal + 1, st);
    if (rc) return rc;
    a.pre_scale = o->d_scal;
  }
  return ts::adam_step(o->ctx, params, grads, grad_dt, m, v, params_bf16, o->d_items, o->nitems, a, st);
}
int ts_optim_step(ts_optim* o, float* params, const float* grads, float* m, float* v, void* params_bf16, float lr,
                  float beta1, float beta2, float eps, int32_t step, float global_clip, float clipnorm,
                  int32_t fuse_global_clip, void* stream) {
  return optim_step_impl(o, params, grads, TS_F32, m, v, params_bf16, lr, beta1, beta2, eps, step, global_clip, clipnorm, fuse_global_clip, stream);
}
int ts_optim_step_lp(ts_optim* o, float* params, const void* grads_bf16, float* m, float* v, void* params_bf16, float lr,
                     float beta1, float beta2, float eps, int32_t step, float global_clip, float clipnorm,
                     int32_t fuse_global_clip, void* stream) {
  return optim_step_impl(o, params, grads_bf16, TS_BF16, m, v, params_bf16, lr, beta1, beta2, eps, step, global_clip, clipnorm, fuse_global_clip, stream);
}

int ts_cast_f32_to_bf16(ts_ctx* ctx, const float* src, void* dst, int64_t n, void* stream) {
  return ts::cast_f32_to_bf16(reinterpret_cast<Ctx*>(ctx), src, dst, n, (cudaStream_t)stream);
}

int ts_grad_pack_bf16(ts_ctx* ctx, const float* src, void* dst, int64_t n, const float* scale_dev, void* stream) {
  if (!ctx || !src || !dst) return TS_EINVAL;
  return ts::grad_pack_bf16(reinterpret_cast<Ctx*>(ctx), src, dst, n, scale_dev, (cudaStream_t)stream);
}
int ts_grad_unpack_bf16(ts_ctx* ctx, const void* src, float* dst, int64_t n, void* stream) {
  if (!ctx || !src || !dst) return TS_EINVAL;
  return ts::grad_unpack_bf16(reinterpret_cast<Ctx*>(ctx), src, dst, n, (cudaStream_t)stream);
}
int ts_dropout(ts_ctx* ctx_, int dtype, const void* x, void* y, int64_t n, float rate, uint64_t seed, void* stream) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx) return TS_EINVAL;
  TS_REQUIRE(ctx, x && y && n >= 0 && rate >= 0.f && rate < 1.f, TS_EINVAL, "ts_dropout: bad arguments (rate must be in [0, 1))");
  TS_REQUIRE(ctx, dtype == TS_F32 || dtype == TS_BF16, TS_EDTYPE, "ts_dropout: dtype %d", dtype);
  return ts::dropout_apply(ctx, dtype, x, y, n, rate, seed, (cudaStream_t)stream);
}

}  // extern "C"
