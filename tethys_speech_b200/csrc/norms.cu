// K11 LayerNorm fwd/bwd (W:214,216,245,249,253,322,392; V:280,411,415,554,778) and K5 GroupNorm + exact GELU
// fwd/bwd (V:140-196, V:248-249, V:265-266). HBM-bound: 16-byte vector accesses, one warp per row (LN) /
// one thread per 8 channels (GN), statistics in fp32 (LN) or fp64 accumulators (GN, 1.5M-element groups).
#include "ops.cuh"
#include "vec.cuh"

namespace ts {

// ------------------------------------------------------------------------------------------------
// LayerNorm
// ------------------------------------------------------------------------------------------------
template <typename T, int NCH>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     T* __restrict__ y, T* __restrict__ sum_out, float* __restrict__ mean,
                                                     float* __restrict__ rstd, int rows, int cols, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const T* xr = x + (long long)row * cols;
  float v[NCH][8];
  float s = 0.f;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
    if (col < cols) {
      load8<T>(xr + col, v[ch]);
      if (res) {
        float r[8];
        load8<T>(res + (long long)row * cols + col, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[ch][i] += r[i];
      }
      if (sum_out) {
        // the consumer of the sum sees the value rounded to the activation dtype: normalise that value
        store8<T>(sum_out + (long long)row * cols + col, v[ch]);
        round8<T>(v[ch]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[ch][i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[ch][i] = 0.f;
    }
  }
  const float mu = warp_sum(s) / cols;
  float q = 0.f;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
    if (col < cols) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[ch][i] - mu; q += d * d; }
    }
  }
  const float rs = rsqrtf(warp_sum(q) / cols + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
    if (col < cols) {
      float g[8], b[8], o[8];
      load8<float>(gamma + col, g);
      load8<float>(beta + col, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[ch][i] - mu) * rs * g[i] + b[i];
      store8<T>(y + (long long)row * cols + col, o);
    }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const T* __restrict__ dres,
                                                     T* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, int rows, int cols) {
  extern __shared__ float sm[];  // [2][cols]
  float* sg = sm;
  float* sb = sm + cols;
  for (int i = threadIdx.x; i < 2 * cols; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[NCH][8], ab[NCH][8], g[NCH][8];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[ch][i] = 0.f; ab[ch][i] = 0.f; g[ch][i] = 0.f; }
    if (col < cols) load8<float>(gamma + col, g[ch]);
  }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float mu = mean[row], rs = rstd[row];
    float xh[NCH][8], d[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int col = ch * 256 + lane * 8;
      if (col < cols) {
        load8<T>(x + (long long)row * cols + col, xh[ch]);
        load8<T>(dy + (long long)row * cols + col, d[ch]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[ch][i] = (xh[ch][i] - mu) * rs;
          ag[ch][i] += d[ch][i] * xh[ch][i];
          ab[ch][i] += d[ch][i];
          d[ch][i] *= g[ch][i];
          s1 += d[ch][i];
          s2 += d[ch][i] * xh[ch][i];
        }
      }
    }
    s1 = warp_sum(s1) / cols;
    s2 = warp_sum(s2) / cols;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int col = ch * 256 + lane * 8;
      if (col < cols) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rs * (d[ch][i] - s1 - xh[ch][i] * s2);
        if (dres) {
          float r[8];
          load8<T>(dres + (long long)row * cols + col, r);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += r[i];
        }
        store8<T>(dx + (long long)row * cols + col, o);
      }
    }
  }
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
    if (col < cols) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { atomicAdd(&sg[col + i], ag[ch][i]); atomicAdd(&sb[col + i], ab[ch][i]); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    atomicAdd(&dgamma[i], sg[i]);
    atomicAdd(&dbeta[i], sb[i]);
  }
}

template <typename T>
static int ln_fwd_t(Ctx* ctx, const void* x, const void* res, const float* gamma, const float* beta, void* y,
                    void* sum_out, float* mean, float* rstd, int rows, int cols, float eps, cudaStream_t st) {
  const int nch = cdiv(cols, 256);
  dim3 grid(cdiv(rows, 8));
#define LN_FWD_CASE(N)                                                                                         \
  case N:                                                                                                      \
    ln_fwd_kernel<T, N><<<grid, 256, 0, st>>>((const T*)x, (const T*)res, gamma, beta, (T*)y, (T*)sum_out, mean, \
                                              rstd, rows, cols, eps);                                          \
    break;
  switch (nch) {
    LN_FWD_CASE(1) LN_FWD_CASE(2) LN_FWD_CASE(3) LN_FWD_CASE(4) LN_FWD_CASE(5)
    default: return set_err(ctx, TS_EUNSUPPORTED, "layernorm: cols=%d > 1280 unsupported", cols);
  }
#undef LN_FWD_CASE
  TS_LAUNCH_OK(ctx);
  return 0;
}

int layernorm_fwd(Ctx* ctx, int dt, const void* x, const void* res, const float* gamma, const float* beta, void* y,
                  void* sum_out, float* mean, float* rstd, int rows, int cols, float eps, cudaStream_t st) {
  TS_REQUIRE(ctx, cols % 8 == 0 && cols > 0 && rows > 0, TS_ESHAPE, "layernorm: rows=%d cols=%d (cols %% 8 != 0)", rows, cols);
  if (dt == TS_F32) return ln_fwd_t<float>(ctx, x, res, gamma, beta, y, sum_out, mean, rstd, rows, cols, eps, st);
  if (dt == TS_BF16) return ln_fwd_t<bf16>(ctx, x, res, gamma, beta, y, sum_out, mean, rstd, rows, cols, eps, st);
  return set_err(ctx, TS_EDTYPE, "layernorm: dtype %d", dt);
}

template <typename T>
static int ln_bwd_t(Ctx* ctx, const void* dy, const void* x, const float* gamma, const float* mean,
                    const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int cols,
                    cudaStream_t st) {
  const int nch = cdiv(cols, 256);
  const int grid = min(cdiv(rows, 8), ctx->num_sms * 4);
  const size_t smem = 2 * cols * sizeof(float);
#define LN_BWD_CASE(N)                                                                                          \
  case N:                                                                                                       \
    ln_bwd_kernel<T, N><<<grid, 256, smem, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres,  \
                                                 (T*)dx, dgamma, dbeta, rows, cols);                            \
    break;
  switch (nch) {
    LN_BWD_CASE(1) LN_BWD_CASE(2) LN_BWD_CASE(3) LN_BWD_CASE(4) LN_BWD_CASE(5)
    default: return set_err(ctx, TS_EUNSUPPORTED, "layernorm_bwd: cols=%d > 1280 unsupported", cols);
  }
#undef LN_BWD_CASE
  TS_LAUNCH_OK(ctx);
  return 0;
}

int layernorm_bwd(Ctx* ctx, int dt, const void* dy, const void* x, const float* gamma, const float* mean,
                  const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int cols,
                  cudaStream_t st) {
  TS_REQUIRE(ctx, cols % 8 == 0 && cols > 0 && rows > 0, TS_ESHAPE, "layernorm_bwd: rows=%d cols=%d", rows, cols);
  if (dt == TS_F32) return ln_bwd_t<float>(ctx, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, cols, st);
  if (dt == TS_BF16) return ln_bwd_t<bf16>(ctx, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, cols, st);
  return set_err(ctx, TS_EDTYPE, "layernorm_bwd: dtype %d", dt);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (+ exact GELU)
// thread <-> 8 consecutive channels of one row; TPR = C/8 threads per row, 256/TPR rows per block step.
// ------------------------------------------------------------------------------------------------
constexpr int GN_ROWS_PER_BLOCK = 256;  // time rows handled by one block

__global__ void gn_zero_accum(double* accum, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) accum[i] = 0.0;
}

template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, double* __restrict__ accum, int T_,
                                                       int C, int G, long long rpb) {
  __shared__ double sacc[64 * 2];
  const int b = blockIdx.y;
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const int cpg = C / G;
  for (int i = threadIdx.x; i < 2 * G; i += 256) sacc[i] = 0.0;
  __syncthreads();
  float s = 0.f, ss = 0.f;
  const int t0 = blockIdx.x * GN_ROWS_PER_BLOCK;
  const int t1 = min(T_, t0 + GN_ROWS_PER_BLOCK);
  if (tr < rpi) {
    for (int t = t0 + tr; t < t1; t += rpi) {
      float v[8];
      load8<T>(x + ((long long)b * rpb + t) * C + tc * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s += v[i]; ss += v[i] * v[i]; }
    }
    const int g = (tc * 8) / cpg;
    atomicAdd(&sacc[2 * g], (double)s);
    atomicAdd(&sacc[2 * g + 1], (double)ss);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += 256) atomicAdd(&accum[(long long)b * G * 2 + i], sacc[i]);
}

__global__ void gn_finalize_kernel(const double* __restrict__ accum, float* mean, float* rstd, int n, double cnt,
                                   float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mu = accum[2 * i] / cnt;
  double var = accum[2 * i + 1] / cnt - mu * mu;
  if (var < 0) var = 0;
  mean[i] = (float)mu;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

int groupnorm_stats(Ctx* ctx, int dt, const void* x, double* accum, float* mean, float* rstd, int B, int T_, int C,
                    int G, long long rpb, float eps, cudaStream_t st) {
  TS_REQUIRE(ctx, C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && G <= 64 && C % G == 0 && (C / G) % 8 == 0,
             TS_ESHAPE, "groupnorm: unsupported C=%d G=%d", C, G);
  gn_zero_accum<<<cdiv(B * G * 2, 256), 256, 0, st>>>(accum, B * G * 2);
  dim3 grid(cdiv(T_, GN_ROWS_PER_BLOCK), B);
  if (dt == TS_F32) gn_stats_kernel<float><<<grid, 256, 0, st>>>((const float*)x, accum, T_, C, G, rpb);
  else gn_stats_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, accum, T_, C, G, rpb);
  gn_finalize_kernel<<<cdiv(B * G, 256), 256, 0, st>>>(accum, mean, rstd, B * G, (double)T_ * (C / G), eps);
  TS_LAUNCH_OK(ctx);
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256) gn_gelu_fwd_kernel(const T* __restrict__ x, long long x_rpb,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          T* __restrict__ y, long long y_rpb, int y_left, int T_, int C,
                                                          int G) {
  const int b = blockIdx.y;
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  if (tr >= rpi) return;
  const int cpg = C / G, g = (tc * 8) / cpg;
  const float mu = mean[b * G + g], rs = rstd[b * G + g];
  float ga[8], be[8];
  load8<float>(gamma + tc * 8, ga);
  load8<float>(beta + tc * 8, be);
  const long long r0 = (long long)blockIdx.x * GN_ROWS_PER_BLOCK;
  const long long r1 = min(y_rpb, r0 + GN_ROWS_PER_BLOCK);
  for (long long r = r0 + tr; r < r1; r += rpi) {
    const long long t = r - y_left;
    float o[8];
    if (t >= 0 && t < T_) {
      float v[8];
      load8<T>(x + ((long long)b * x_rpb + t) * C + tc * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = gelu_f(ga[i] * ((v[i] - mu) * rs) + be[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
    store8<T>(y + ((long long)b * y_rpb + r) * C + tc * 8, o);
  }
}

int groupnorm_gelu_fwd(Ctx* ctx, int dt, const void* x, long long x_rpb, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, void* y, long long y_rpb, int y_left, int B, int T_,
                       int C, int G, cudaStream_t st) {
  dim3 grid(cdiv(y_rpb, GN_ROWS_PER_BLOCK), B);
  if (dt == TS_F32)
    gn_gelu_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, x_rpb, mean, rstd, gamma, beta, (float*)y, y_rpb,
                                                    y_left, T_, C, G);
  else
    gn_gelu_fwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, x_rpb, mean, rstd, gamma, beta, (bf16*)y, y_rpb,
                                                   y_left, T_, C, G);
  TS_LAUNCH_OK(ctx);
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256) gn_gelu_bwd1_kernel(const T* __restrict__ da, long long da_rpb, Col2imSrc col,
                                                           const T* __restrict__ x, long long x_rpb,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           T* __restrict__ dx, long long dx_rpb, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, double* __restrict__ accum, int T_,
                                                           int C, int G) {
  extern __shared__ float smf[];  // [2][C] floats then [2*G] doubles (8-byte aligned by construction)
  float* sg = smf;
  float* sb = smf + C;
  double* sacc = reinterpret_cast<double*>(smf + 2 * C);
  const int b = blockIdx.y;
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const int cpg = C / G, g = (tc * 8) / cpg;
  for (int i = threadIdx.x; i < 2 * C; i += 256) smf[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * G; i += 256) sacc[i] = 0.0;
  __syncthreads();
  if (tr < rpi) {
    const float mu = mean[b * G + g], rs = rstd[b * G + g];
    float ga[8], be[8], ag[8], ab[8];
    load8<float>(gamma + tc * 8, ga);
    load8<float>(beta + tc * 8, be);
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
    float s1 = 0.f, s2 = 0.f;
    const int t0 = blockIdx.x * GN_ROWS_PER_BLOCK, t1 = min(T_, t0 + GN_ROWS_PER_BLOCK);
    const T* dcol = reinterpret_cast<const T*>(col.dcol);
    for (int t = t0 + tr; t < t1; t += rpi) {
      float d[8];
      if (dcol) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = 0.f;
        for (int j = 0; j < col.k; ++j) {
          const int q = t + col.left - j;
          if (q >= 0 && (q % col.s) == 0 && (q / col.s) < col.t_next) {
            float v[8];
            load8<T>(dcol + ((long long)b * col.rows_per_batch + q / col.s) * ((long long)col.k * C) + (long long)j * C + tc * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] += v[i];
          }
        }
      } else {
        load8<T>(da + ((long long)b * da_rpb + t) * C + tc * 8, d);
      }
      float v[8], o[8];
      load8<T>(x + ((long long)b * x_rpb + t) * C + tc * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (v[i] - mu) * rs;
        const float u = ga[i] * xh + be[i];
        const float dact = d[i] * gelu_grad_f(u);
        o[i] = dact;
        ag[i] += dact * xh;
        ab[i] += dact;
        s1 += dact * ga[i];
        s2 += dact * ga[i] * xh;
      }
      store8<T>(dx + ((long long)b * dx_rpb + t) * C + tc * 8, o);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { atomicAdd(&sg[tc * 8 + i], ag[i]); atomicAdd(&sb[tc * 8 + i], ab[i]); }
    atomicAdd(&sacc[2 * g], (double)s1);
    atomicAdd(&sacc[2 * g + 1], (double)s2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) { atomicAdd(&dgamma[i], sg[i]); atomicAdd(&dbeta[i], sb[i]); }
  for (int i = threadIdx.x; i < 2 * G; i += 256) atomicAdd(&accum[(long long)b * G * 2 + i], sacc[i]);
}

template <typename T>
__global__ void __launch_bounds__(256) gn_gelu_bwd2_kernel(const T* __restrict__ x, long long x_rpb,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, T* __restrict__ dx,
                                                           long long dx_rpb, const double* __restrict__ accum, int T_, int C,
                                                           int G) {
  const int b = blockIdx.y;
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  if (tr >= rpi) return;
  const int cpg = C / G, g = (tc * 8) / cpg;
  const float mu = mean[b * G + g], rs = rstd[b * G + g];
  const double n = (double)T_ * cpg;
  const float m1 = (float)(accum[((long long)b * G + g) * 2] / n), m2 = (float)(accum[((long long)b * G + g) * 2 + 1] / n);
  float ga[8];
  load8<float>(gamma + tc * 8, ga);
  const long long r0 = (long long)blockIdx.x * GN_ROWS_PER_BLOCK, r1 = min(dx_rpb, r0 + GN_ROWS_PER_BLOCK);
  for (long long t = r0 + tr; t < r1; t += rpi) {
    float o[8];
    T* p = dx + ((long long)b * dx_rpb + t) * C + tc * 8;
    if (t < T_) {
      float v[8], d[8];
      load8<T>(x + ((long long)b * x_rpb + t) * C + tc * 8, v);
      load8<T>(p, d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (v[i] - mu) * rs;
        o[i] = rs * (d[i] * ga[i] - m1 - xh * m2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
    store8<T>(p, o);
  }
}

int groupnorm_gelu_bwd(Ctx* ctx, int dt, const void* da, long long da_rpb, const Col2imSrc* col, const void* x,
                       long long x_rpb, const float* mean, const float* rstd, const float* gamma, const float* beta,
                       void* dx, long long dx_rpb, float* dgamma, float* dbeta, double* accum, int B, int T_, int C, int G,
                       cudaStream_t st) {
  Col2imSrc c;
  if (col) c = *col; else { c.dcol = nullptr; c.rows_per_batch = 0; c.t_next = 0; c.k = 0; c.s = 1; c.left = 0; }
  gn_zero_accum<<<cdiv(B * G * 2, 256), 256, 0, st>>>(accum, B * G * 2);
  const size_t smem = 2 * C * sizeof(float) + 2 * G * sizeof(double);
  dim3 g1(cdiv(T_, GN_ROWS_PER_BLOCK), B), g2(cdiv(dx_rpb, GN_ROWS_PER_BLOCK), B);
  if (dt == TS_F32) {
    gn_gelu_bwd1_kernel<float><<<g1, 256, smem, st>>>((const float*)da, da_rpb, c, (const float*)x, x_rpb, mean, rstd,
                                                      gamma, beta, (float*)dx, dx_rpb, dgamma, dbeta, accum, T_, C, G);
    gn_gelu_bwd2_kernel<float><<<g2, 256, 0, st>>>((const float*)x, x_rpb, mean, rstd, gamma, (float*)dx, dx_rpb, accum,
                                                   T_, C, G);
  } else {
    gn_gelu_bwd1_kernel<bf16><<<g1, 256, smem, st>>>((const bf16*)da, da_rpb, c, (const bf16*)x, x_rpb, mean, rstd, gamma,
                                                     beta, (bf16*)dx, dx_rpb, dgamma, dbeta, accum, T_, C, G);
    gn_gelu_bwd2_kernel<bf16><<<g2, 256, 0, st>>>((const bf16*)x, x_rpb, mean, rstd, gamma, (bf16*)dx, dx_rpb, accum, T_,
                                                  C, G);
  }
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
