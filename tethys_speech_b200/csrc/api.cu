// C-ABI: context lifecycle, error reporting, GEMM dispatch.
#include <stdarg.h>
#include "common.cuh"

namespace ts {

struct CtxHolder { Ctx c; };

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
  if (d->force_engine == 1) return gemm_simt(ctx, d, st);
  if (d->force_engine == 2) return gemm_tc(ctx, d, st);
  if (d->in_dtype == TS_BF16 && gemm_tc_supported(d)) return gemm_tc(ctx, d, st);
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
  if (cudaMalloc(&c->d_watchdog, sizeof(int)) != cudaSuccess) { delete c; return TS_ECUDA; }
  cudaMemset(c->d_watchdog, 0, sizeof(int));
  *out = reinterpret_cast<ts_ctx*>(c);
  return TS_OK;
}

void ts_destroy(ts_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return;
  ts::tmap_cache_free(c);
  if (c->d_watchdog) cudaFree(c->d_watchdog);
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

int ts_gemm(ts_ctx* ctx, const ts_gemm_desc* d, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !d) return TS_EINVAL;
  return ts::gemm(c, d, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
