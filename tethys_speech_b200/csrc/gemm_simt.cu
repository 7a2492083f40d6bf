// K9 — CUDA-core GEMM engine: true fp32 FMA accumulation for the 1e-5 parity mode, and the path for
// operands whose strides/alignment the TMA engine cannot describe. Same descriptor and epilogue as
// gemm_tc.cu (include/tethys.h: ts_gemm_desc).
#include "common.cuh"

namespace ts {

struct SimtParams {
  const void* a; const void* b; void* c; void* c_pre; const void* res; const float* bias; const void* aux;
  long long lda, ldb, ldc, ldr, ld_aux;
  long long a_bs1, a_bs2, b_bs1, b_bs2, c_bs1, c_bs2, r_bs1, r_bs2, bias_bs1;
  int m, n, k, nb1;
  int a_major, b_major;
  float alpha; int act, accumulate;
  uint32_t drop_thr; float inv_keep; unsigned long long seed;
  const unsigned long long* salt;
};

constexpr int SB_M = 64, SB_N = 64, SB_K = 16;

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p) {
  ts::pdl_enter();
  __shared__ float As[SB_K][SB_M + 4];
  __shared__ float Bs[SB_K][SB_N + 4];
  const int b1 = blockIdx.z % p.nb1, b2 = blockIdx.z / p.nb1;
  const TI* A = reinterpret_cast<const TI*>(p.a) + b1 * p.a_bs1 + b2 * p.a_bs2;
  const TI* B = reinterpret_cast<const TI*>(p.b) + b1 * p.b_bs1 + b2 * p.b_bs2;
  const int m0 = blockIdx.x * SB_M, n0 = blockIdx.y * SB_N;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 micro tile each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.k; k0 += SB_K) {
    // load A tile (64 x 16) and B tile (64 x 16): 1024 elements each, 4 per thread
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = threadIdx.x + it * 256;
      int mm, kk;
      if (p.a_major == 0) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < p.m && gk < p.k)
        v = to_f<TI>(p.a_major == 0 ? A[(long long)gm * p.lda + gk] : A[(long long)gk * p.lda + gm]);
      As[kk][mm] = v;
      int nn, kb;
      if (p.b_major == 0) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
      const int gn = n0 + nn, gk2 = k0 + kb;
      float w = 0.f;
      if (gn < p.n && gk2 < p.k)
        w = to_f<TI>(p.b_major == 0 ? B[(long long)gn * p.ldb + gk2] : B[(long long)gk2 * p.ldb + gn]);
      Bs[kb][nn] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SB_K; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  TO* C = reinterpret_cast<TO*>(p.c) + b1 * p.c_bs1 + b2 * p.c_bs2;
  TO* CP = p.c_pre ? reinterpret_cast<TO*>(p.c_pre) + b1 * p.c_bs1 + b2 * p.c_bs2 : nullptr;
  const TO* R = p.res ? reinterpret_cast<const TO*>(p.res) + b1 * p.r_bs1 + b2 * p.r_bs2 : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= p.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= p.n) continue;
      float v = acc[i][j] * p.alpha;
      if (p.bias) v += p.bias[b1 * p.bias_bs1 + gn];
      if (CP) CP[(long long)gm * p.ldc + gn] = from_f<TO>(v);
      if (p.act == 1) v = gelu_f(v);
      else if (p.act == 2) v *= gelu_grad_f(to_f<TO>(reinterpret_cast<const TO*>(p.aux)[b1 * p.c_bs1 + b2 * p.c_bs2 + (long long)gm * p.ld_aux + gn]));
      if (p.drop_thr)
        v *= dropout_scale(salted_seed(p.seed, p.salt), (unsigned long long)(b1 * p.c_bs1 + b2 * p.c_bs2 + (long long)gm * p.ldc + gn), p.drop_thr, p.inv_keep);
      if (R) v += to_f<TO>(R[(long long)gm * p.ldr + gn]);
      if (p.accumulate) v += to_f<TO>(C[(long long)gm * p.ldc + gn]);
      C[(long long)gm * p.ldc + gn] = from_f<TO>(v);
    }
  }
}

int gemm_simt(Ctx* ctx, const ts_gemm_desc* d, cudaStream_t st) {
  SimtParams p;
  p.a = d->a; p.b = d->b; p.c = d->c; p.c_pre = d->c_preact; p.res = d->residual; p.bias = d->bias;
  p.lda = d->lda; p.ldb = d->ldb; p.ldc = d->ldc; p.ldr = d->ldr; p.aux = d->act_aux; p.ld_aux = d->ld_aux;
  p.a_bs1 = d->a_bs1; p.a_bs2 = d->a_bs2; p.b_bs1 = d->b_bs1; p.b_bs2 = d->b_bs2;
  p.c_bs1 = d->c_bs1; p.c_bs2 = d->c_bs2; p.r_bs1 = d->r_bs1; p.r_bs2 = d->r_bs2; p.bias_bs1 = d->bias_bs1;
  p.m = d->m; p.n = d->n; p.k = d->k;
  const int nb1 = d->batch1 > 0 ? d->batch1 : 1, nb2 = d->batch2 > 0 ? d->batch2 : 1;
  p.nb1 = nb1;
  p.a_major = d->a_major; p.b_major = d->b_major;
  p.alpha = d->alpha; p.act = d->act; p.accumulate = d->accumulate;
  p.drop_thr = 0; p.inv_keep = 1.f; p.seed = d->seed; p.salt = ctx->d_state;
  if (d->drop > 0.f) {
    double t = (double)d->drop * 4294967296.0;
    p.drop_thr = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
    p.inv_keep = 1.f / (1.f - d->drop);
  }
  TS_REQUIRE(ctx, d->m > 0 && d->n > 0 && d->k > 0, TS_ESHAPE, "gemm: empty problem m=%d n=%d k=%d", d->m, d->n, d->k);
  dim3 grid(cdiv(d->m, SB_M), cdiv(d->n, SB_N), nb1 * nb2);
  if (d->in_dtype == TS_F32 && d->out_dtype == TS_F32) ts::launch_k(gemm_simt_kernel<float, float>, grid, 256, 0, st, p);
  else if (d->in_dtype == TS_BF16 && d->out_dtype == TS_BF16) ts::launch_k(gemm_simt_kernel<bf16, bf16>, grid, 256, 0, st, p);
  else if (d->in_dtype == TS_BF16 && d->out_dtype == TS_F32) ts::launch_k(gemm_simt_kernel<bf16, float>, grid, 256, 0, st, p);
  else return set_err(ctx, TS_EDTYPE, "gemm_simt: unsupported dtype combination in=%d out=%d", d->in_dtype, d->out_dtype);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
