// Standalone GPU self-test of the tcgen05 GEMM engine (no Python/torch): compares gemm_tc against a
// double-precision CPU reference on bf16-rounded inputs for every operand-major combination, ragged
// shapes, batches, overlapping (conv-window) rows and the epilogue options, and times the big shapes.
//   build: see tools/Makefile ; run on a B200: ./tools/selftest_gemm
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../include/tethys.h"

static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

struct Case {
  const char* name;
  int m, n, k, amaj, bmaj, nb1, nb2;
  long long lda, ldb;  // 0 => dense
  int out_f32, bias, act, res, accumulate, preact;
  float alpha;
  float drop;  // timing cases only (the mask is checked by the parity tests, not here)
};

static void* g_flush = nullptr;           // 512 MB scratch: written between launches to time a kernel with a cold L2
static const size_t kFlushBytes = 512ull << 20;

static int run_case(ts_ctx* ctx, const Case& cs, int engine, bool timing) {
  const int m = cs.m, n = cs.n, k = cs.k, nb = cs.nb1 * cs.nb2;
  const long long a_rows = cs.amaj == 0 ? m : k, a_cols = cs.amaj == 0 ? k : m;
  const long long b_rows = cs.bmaj == 0 ? n : k, b_cols = cs.bmaj == 0 ? k : n;
  const long long lda = cs.lda ? cs.lda : a_cols, ldb = cs.ldb ? cs.ldb : b_cols;
  // storage per batch (rows may overlap when ld < cols)
  const long long a_sz = (a_rows - 1) * lda + a_cols, b_sz = (b_rows - 1) * ldb + b_cols;
  const long long a_bs = ((a_sz + 7) / 8) * 8, b_bs = ((b_sz + 7) / 8) * 8;
  const long long ldc = n, c_bs = (long long)m * n;
  std::vector<float> ha(a_bs * nb), hb(b_bs * nb), hres(c_bs * nb), hc0(c_bs * nb), hbias(n);
  for (auto& v : ha) v = bf16_round(frand());
  for (auto& v : hb) v = bf16_round(frand());
  for (auto& v : hres) v = bf16_round(frand());
  for (auto& v : hc0) v = bf16_round(frand());
  for (auto& v : hbias) v = frand();
  std::vector<__nv_bfloat16> ha16(ha.size()), hb16(hb.size());
  for (size_t i = 0; i < ha.size(); ++i) ha16[i] = __float2bfloat16_rn(ha[i]);
  for (size_t i = 0; i < hb.size(); ++i) hb16[i] = __float2bfloat16_rn(hb[i]);
  const size_t osz = cs.out_f32 ? 4 : 2;
  void *da, *db, *dc, *dres, *dpre; float* dbias;
  cudaMalloc(&da, ha16.size() * 2); cudaMalloc(&db, hb16.size() * 2);
  cudaMalloc(&dc, c_bs * nb * osz); cudaMalloc(&dres, c_bs * nb * osz); cudaMalloc(&dpre, c_bs * nb * osz);
  cudaMalloc(&dbias, n * 4);
  cudaMemcpy(da, ha16.data(), ha16.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb16.data(), hb16.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dbias, hbias.data(), n * 4, cudaMemcpyHostToDevice);
  if (cs.out_f32) {
    cudaMemcpy(dres, hres.data(), c_bs * nb * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, hc0.data(), c_bs * nb * 4, cudaMemcpyHostToDevice);
  } else {
    std::vector<__nv_bfloat16> t(c_bs * nb);
    for (size_t i = 0; i < t.size(); ++i) t[i] = __float2bfloat16_rn(hres[i]);
    cudaMemcpy(dres, t.data(), t.size() * 2, cudaMemcpyHostToDevice);
    for (size_t i = 0; i < t.size(); ++i) t[i] = __float2bfloat16_rn(hc0[i]);
    cudaMemcpy(dc, t.data(), t.size() * 2, cudaMemcpyHostToDevice);
  }
  ts_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = da; d.b = db; d.c = dc; d.m = m; d.n = n; d.k = k; d.a_major = cs.amaj; d.b_major = cs.bmaj;
  d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.batch1 = cs.nb1; d.batch2 = cs.nb2;
  d.a_bs1 = a_bs; d.a_bs2 = a_bs * cs.nb1; d.b_bs1 = b_bs; d.b_bs2 = b_bs * cs.nb1; d.c_bs1 = c_bs; d.c_bs2 = c_bs * cs.nb1;
  d.in_dtype = TS_BF16; d.out_dtype = cs.out_f32 ? TS_F32 : TS_BF16; d.alpha = cs.alpha;
  d.bias = cs.bias ? dbias : nullptr; d.act = cs.act; d.residual = cs.res ? dres : nullptr; d.ldr = ldc;
  d.r_bs1 = c_bs; d.r_bs2 = c_bs * cs.nb1; d.accumulate = cs.accumulate; d.c_preact = cs.preact ? dpre : nullptr;
  d.force_engine = engine;
  d.drop = timing ? cs.drop : 0.f; d.seed = 0x1234;
  int rc = ts_gemm(ctx, &d, 0);
  if (rc) { printf("  [%s] ts_gemm rc=%d: %s\n", cs.name, rc, ts_last_error(ctx)); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  [%s] CUDA error: %s\n", cs.name, cudaGetErrorString(e)); return 2; }
  rc = ts_watchdog_check(ctx);
  if (rc) { printf("  [%s] watchdog: %s\n", cs.name, ts_last_error(ctx)); return 3; }

  int bad = 0;
  if (!timing) {
    std::vector<float> hc(c_bs * nb), hp(c_bs * nb);
    if (cs.out_f32) {
      cudaMemcpy(hc.data(), dc, hc.size() * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(hp.data(), dpre, hp.size() * 4, cudaMemcpyDeviceToHost);
    } else {
      std::vector<__nv_bfloat16> t(c_bs * nb);
      cudaMemcpy(t.data(), dc, t.size() * 2, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < t.size(); ++i) hc[i] = __bfloat162float(t[i]);
      cudaMemcpy(t.data(), dpre, t.size() * 2, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < t.size(); ++i) hp[i] = __bfloat162float(t[i]);
    }
    double max_err = 0, max_ref = 0;
    for (int z = 0; z < nb; ++z)
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
          double acc = 0;
          const float* A = ha.data() + z * a_bs;
          const float* B = hb.data() + z * b_bs;
          for (int kk = 0; kk < k; ++kk) {
            const double av = cs.amaj == 0 ? A[i * lda + kk] : A[kk * lda + i];
            const double bv = cs.bmaj == 0 ? B[j * ldb + kk] : B[kk * ldb + j];
            acc += av * bv;
          }
          double v = acc * cs.alpha;
          if (cs.bias) v += hbias[j];
          const double pre = v;
          if (cs.act) v = 0.5 * v * (1.0 + erf(v / sqrt(2.0)));
          if (cs.res) v += hres[z * c_bs + (long long)i * n + j];
          if (cs.accumulate) v += hc0[z * c_bs + (long long)i * n + j];
          const double got = hc[z * c_bs + (long long)i * n + j];
          double err = fabs(got - v);
          if (cs.preact) err = fmax(err, fabs((double)hp[z * c_bs + (long long)i * n + j] - pre));
          const double tol = (cs.out_f32 ? 2e-4 : 1.2e-2) * (1.0 + fmax(fabs(v), cs.preact ? fabs(pre) : 0.0));
          if (err > tol) {
            if (bad < 5) printf("    mismatch z=%d i=%d j=%d got=%g ref=%g\n", z, i, j, got, v);
            ++bad;
          }
          max_err = fmax(max_err, err); max_ref = fmax(max_ref, fabs(v));
        }
    printf("  [%-28s] eng=%d m=%d n=%d k=%d maj=%d%d nb=%dx%d lda=%lld ldb=%lld  max_err=%.3e (max_ref %.3g) %s\n",
           cs.name, engine, m, n, k, cs.amaj, cs.bmaj, cs.nb1, cs.nb2, lda, ldb, max_err, max_ref, bad ? "FAIL" : "ok");
  } else {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) ts_gemm(ctx, &d, 0);
    cudaEventRecord(e0);
    const int iters = 20;
    for (int i = 0; i < iters; ++i) ts_gemm(ctx, &d, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double tf = 2.0 * m * n * (double)k * nb / (ms * 1e-3) / 1e12;
    // cold: L2 flushed before every launch (what the kernel sees inside a train step, whose activations far exceed L2)
    if (!g_flush) cudaMalloc(&g_flush, kFlushBytes);
    float cold_ms = 0;
    const int citers = 8;
    for (int i = 0; i < citers; ++i) {
      cudaMemsetAsync(g_flush, i, kFlushBytes);
      cudaEventRecord(e0);
      ts_gemm(ctx, &d, 0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float t = 0;
      cudaEventElapsedTime(&t, e0, e1);
      cold_ms += t;
    }
    cold_ms /= citers;
    const double tfc = 2.0 * m * n * (double)k * nb / (cold_ms * 1e-3) / 1e12;
    printf("  [time %-22s] eng=%d m=%d n=%d k=%d maj=%d%d nb=%d : warm %.1f us %.1f TFLOP/s | cold-L2 %.1f us %.1f TFLOP/s\n", cs.name, engine,
           m, n, k, cs.amaj, cs.bmaj, nb, ms * 1e3, tf, cold_ms * 1e3, tfc);
  }
  cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dres); cudaFree(dpre); cudaFree(dbias);
  return bad ? 4 : 0;
}

int main(int argc, char** argv) {
  ts_ctx* ctx = nullptr;
  int rc = ts_create(0, &ctx);
  if (rc) { printf("ts_create failed rc=%d\n", rc); return 1; }
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  const bool time_only = argc > 1 && !strcmp(argv[1], "time");   // timing section only (tile-choice experiments under TETHYS_GEMM_* switches)
  if (argc > 1 && !strcmp(argv[1], "prof")) {  // the workload's dominant GEMM only (for ncu): FFN fc1 + bias + GELU + pre-activation copy
    Case c = {"ffn1+gelu+preact", 6000, 3072, 768, 0, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0, 1, 1.f};
    run_case(ctx, c, 2, true);
    ts_destroy(ctx);
    return 0;
  }
  std::vector<Case> cases = {
      // name, m, n, k, amaj, bmaj, nb1, nb2, lda, ldb, out_f32, bias, act, res, acc, preact, alpha
      {"kk_basic", 128, 128, 64, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"kk_k256", 128, 128, 256, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"kk_multi_tile", 384, 256, 320, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"kn_dense_fwd", 256, 192, 128, 0, 1, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"nk", 256, 128, 192, 1, 0, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"nn_wgrad", 256, 192, 200, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"ragged_kk", 200, 72, 104, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"ragged_kn", 150, 328, 88, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
      {"ragged_nn", 136, 200, 77, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"bn64_pv", 100, 64, 136, 0, 1, 3, 2, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
      {"bn256", 256, 1024, 128, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
      {"batched_qk", 100, 100, 64, 0, 0, 4, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0.125f},
      {"epi_bias_gelu_res", 256, 256, 128, 0, 1, 1, 1, 0, 0, 0, 1, 1, 1, 0, 1, 1.f},
      {"epi_accumulate", 128, 256, 192, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 1.f},
      {"conv_window_rows", 120, 128, 192, 0, 1, 2, 1, 128, 0, 0, 1, 0, 0, 0, 0, 1.f},
      {"conv_window_wgrad", 192, 128, 120, 1, 1, 1, 1, 128, 0, 1, 0, 0, 0, 0, 0, 1.f},
      {"persist_many_tiles", 2560, 2304, 128, 0, 1, 1, 1, 0, 0, 0, 1, 1, 1, 0, 1, 1.f},
      {"persist_many_f32", 1920, 1280, 192, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 0.5f},
      {"splitk_wgrad", 256, 256, 4096, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 1.f},
      {"splitk_ragged", 200, 136, 3000, 1, 1, 2, 1, 0, 0, 1, 0, 0, 0, 1, 0, 1.f},
      {"odd_n_bf16", 150, 75, 72, 0, 0, 2, 1, 0, 0, 0, 1, 1, 1, 0, 1, 1.f},
      {"batched_many", 300, 300, 64, 0, 0, 12, 8, 0, 0, 0, 0, 0, 0, 0, 0, 0.125f},
  };
  int fails = 0;
  if (!time_only) {
  printf("== correctness: tcgen05 engine (2) vs fp64 CPU reference ==\n");
  for (auto& c : cases) fails += run_case(ctx, c, 2, false) != 0;
  printf("== correctness: tcgen05 CTA-pair engine (3: cta_group::2, 256-row tiles) ==\n");
  for (auto& c : cases)
    if (c.m > 128 && c.n > 64) fails += run_case(ctx, c, 3, false) != 0;
  {
    std::vector<Case> more = {
        {"pair_kk_rag_m", 6000 / 4, 768, 320, 0, 0, 1, 1, 0, 0, 0, 1, 0, 1, 0, 0, 1.f},
        {"pair_kn_rag_m", 300, 3072 / 4, 256, 0, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0, 1, 1.f},
        {"pair_nn_splitk", 768, 512, 6000, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 1.f},
        {"pair_nk_batched", 400, 256, 192, 1, 0, 3, 2, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
    };
    for (auto& c : more) fails += run_case(ctx, c, 3, false) != 0;
  }
  if (getenv("TETHYS_SELFTEST_MC")) {
    printf("== correctness: 4-CTA cluster engine (4: two CTA pairs, B multicast) ==\n");
    for (auto& c : cases)
      if (c.m > 256 && c.n > 128) fails += run_case(ctx, c, 4, false) != 0;
    std::vector<Case> more = {
        {"mc_kk_rag_m", 6000 / 4, 768, 320, 0, 0, 1, 1, 0, 0, 0, 1, 0, 1, 0, 0, 1.f},
        {"mc_kn_rag_m", 700, 3072 / 4, 256, 0, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0, 1, 1.f},
        {"mc_nn_splitk", 768, 512, 6000, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 1.f},
        {"mc_nk_batched", 600, 256, 192, 1, 0, 3, 2, 0, 0, 1, 0, 0, 0, 0, 0, 1.f},
    };
    for (auto& c : more) fails += run_case(ctx, c, 4, false) != 0;
  }
  printf("== correctness: CUDA-core engine (1) ==\n");
  for (size_t i = 0; i < cases.size(); i += 3) fails += run_case(ctx, cases[i], 1, false) != 0;
  }
  if (!quick) {
    printf("== timing ==\n");
    std::vector<Case> tcases = {
        {"ffn1 6000x3072x768", 6000, 3072, 768, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 1.f},
        {"ffn2 6000x768x3072", 6000, 768, 3072, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 1.f},
        {"dgrad 6000x768x3072", 6000, 768, 3072, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
        {"ffn1+gelu+preact", 6000, 3072, 768, 0, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0, 1, 1.f},
        {"ffn1+gelu+preact+drop", 6000, 3072, 768, 0, 1, 1, 1, 0, 0, 0, 1, 1, 0, 0, 1, 1.f, 0.1f},
        {"ffn2+res+drop", 6000, 768, 3072, 0, 1, 1, 1, 0, 0, 0, 1, 0, 1, 0, 0, 1.f, 0.1f},
        {"qkv 6000x2304x768", 6000, 2304, 768, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 1.f},
        {"out 6000x768x768 +res", 6000, 768, 768, 0, 1, 1, 1, 0, 0, 0, 1, 0, 1, 0, 0, 1.f},
        {"wgrad 768x3072x6000", 768, 3072, 6000, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 1.f},
        {"wgrad 768x768x6000", 768, 768, 6000, 1, 1, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 1.f},
        {"convwgrad 1536x512x96000", 1536, 512, 96000, 1, 1, 1, 1, 1024, 0, 1, 0, 0, 0, 1, 0, 1.f},
        {"conv1 192000x512x1536", 192000, 512, 1536, 0, 1, 1, 1, 1024, 0, 0, 0, 0, 0, 0, 0, 1.f},
        {"sq 8192^3", 8192, 8192, 8192, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
        {"sq 8192^3 kn", 8192, 8192, 8192, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
        {"wb dgrad 6000x512x512", 6000, 512, 512, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
        {"wb fc1 6000x2048x512", 6000, 2048, 512, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 1.f},
        {"wb dgrad 6000x512x2048", 6000, 512, 2048, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
        {"wb qkv 6000x1536x512", 6000, 1536, 512, 0, 1, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 1.f},
        {"dec 400x768x768", 400, 768, 768, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
        {"dec 400x768x3072", 400, 768, 3072, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1.f},
        {"qk 750x750x64 x96", 750, 750, 64, 0, 0, 12, 8, 0, 0, 0, 0, 0, 0, 0, 0, 0.125f},
        {"pv 750x64x750 x96", 750, 64, 750, 0, 1, 12, 8, 752, 0, 0, 0, 0, 0, 0, 0, 1.f},
    };
    // engine 2 = automatic tile choice (1-CTA or CTA-pair tiles; TETHYS_GEMM_CTAS=1 pins the former), 3 = CTA pair forced
    for (auto& c : tcases) {
      run_case(ctx, c, 2, true);
      if (c.m > 128 && c.n > 64) run_case(ctx, c, 3, true);
      if (getenv("TETHYS_SELFTEST_MC") && c.m > 256 && c.n > 128) run_case(ctx, c, 4, true);
    }
    if (!time_only) run_case(ctx, tcases[0], 1, true);
  }
  printf("== %s (%d failing cases) ==\n", fails ? "SELFTEST FAILED" : "SELFTEST PASSED", fails);
  ts_destroy(ctx);
  return fails ? 1 : 0;
}
