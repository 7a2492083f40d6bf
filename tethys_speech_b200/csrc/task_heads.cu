// Kernels of the two fine-tuning heads that sit on the Wav2Vec2 trunk:
//   Wav2Vec2ForCTC (V:940-1001)                   : dropout -> lm_head -> mean sparse CE against class 0 on every frame
//   Wav2Vec2ForSequenceClassification (V:1004-1070): mean over time -> Dense+tanh -> dropout -> Dense -> mean sparse CE
// The Dense layers go through ts::gemm; here are the row-wise CE (forward + gradient in one pass), the time pooling and
// the tanh/dropout pair. All are tiny next to the trunk (<= B*T*32 logits), one warp per row, HBM/latency-bound.
#include "ops.cuh"
#include "vec.cuh"

namespace ts {

// loss_sum += sum_r (logsumexp(x_r) - x_r[label_r]);  d[r, j] = (softmax(x_r)[j] - [j == label_r]) * grad_scale
template <typename TD>
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ x, long long ld, const int* __restrict__ labels,
                                                      TD* __restrict__ d, long long ld_d, float* __restrict__ loss_sum, int R,
                                                      int V, float grad_scale) {
  ts::pdl_enter();
  __shared__ float part[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  float loss = 0.f;
  if (r < R) {
    const float* xr = x + (long long)r * ld;
    int lab = labels ? labels[r] : 0;
    float mx = -INFINITY;
    for (int j = lane; j < V; j += 32) mx = fmaxf(mx, xr[j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < V; j += 32) s += expf(xr[j] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    const bool lab_ok = lab >= 0 && lab < V;   // out-of-range label: TF returns NaN loss on GPU; here the row contributes lse only
    loss = lse - (lab_ok ? xr[lab] : 0.f);
    if (d) {
      const float inv = 1.f / s;
      for (int j = lane; j < V; j += 32) {
        const float p = expf(xr[j] - mx) * inv;
        d[(long long)r * ld_d + j] = from_f<TD>((p - (j == lab ? 1.f : 0.f)) * grad_scale);
      }
    }
  }
  if (lane == 0) part[warp] = (r < R) ? loss : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w];
    atomicAdd(loss_sum, t);
  }
}

int ce_rows_fwd_bwd(Ctx* ctx, int dt, const float* logits, long long ld, const int* labels, void* dlogits, long long ld_d,
                    float* loss_sum, int R, int V, float grad_scale, cudaStream_t st) {
  if (R <= 0) return 0;
  const int grid = cdiv(R, 8);
  if (dt == TS_F32) ts::launch_k(ce_rows_kernel<float>, grid, 256, 0, st, logits, ld, labels, (float*)dlogits, ld_d, loss_sum, R, V, grad_scale);
  else if (dt == TS_BF16) ts::launch_k(ce_rows_kernel<bf16>, grid, 256, 0, st, logits, ld, labels, (bf16*)dlogits, ld_d, loss_sum, R, V, grad_scale);
  else return set_err(ctx, TS_EDTYPE, "ce_rows: dtype %d", dt);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// pooled[b, c] = mean_t x[b, t, c]   (tf.reduce_mean(hidden_states, axis=1), V:1043). One thread per (b, c) column,
// consecutive threads on consecutive channels (coalesced); fp32 accumulation in time order.
template <typename T>
__global__ void __launch_bounds__(256) mean_pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int Tn, int C) {
  ts::pdl_enter();
  const int c = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (c >= C) return;
  const T* p = x + (long long)b * Tn * C + c;
  float s = 0.f;
  for (int t = 0; t < Tn; ++t) s += to_f<T>(p[(long long)t * C]);
  y[(long long)b * C + c] = from_f<T>(s / (float)Tn);
}
// dx[b, t, c] = dy[b, c] / T
template <typename T>
__global__ void __launch_bounds__(256) mean_pool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int Tn, int C) {
  ts::pdl_enter();
  const long long n = (long long)B * Tn * C;
  const float inv = 1.f / (float)Tn;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C);
    const int b = (int)(i / ((long long)Tn * C));
    dx[i] = from_f<T>(to_f<T>(dy[(long long)b * C + c]) * inv);
  }
}

int mean_pool_fwd(Ctx* ctx, int dt, const void* x, void* y, int B, int Tn, int C, cudaStream_t st) {
  dim3 grid(cdiv(C, 256), B);
  if (dt == TS_F32) ts::launch_k(mean_pool_fwd_kernel<float>, grid, 256, 0, st, (const float*)x, (float*)y, B, Tn, C);
  else if (dt == TS_BF16) ts::launch_k(mean_pool_fwd_kernel<bf16>, grid, 256, 0, st, (const bf16*)x, (bf16*)y, B, Tn, C);
  else return set_err(ctx, TS_EDTYPE, "mean_pool: dtype %d", dt);
  TS_LAUNCH_OK(ctx);
  return 0;
}
int mean_pool_bwd(Ctx* ctx, int dt, const void* dy, void* dx, int B, int Tn, int C, cudaStream_t st) {
  const long long n = (long long)B * Tn * C;
  long long g = (n + 255) / 256;
  const long long cap = (long long)ctx->num_sms * 16;
  const int grid = (int)(g > cap ? cap : g);
  if (dt == TS_F32) ts::launch_k(mean_pool_bwd_kernel<float>, grid, 256, 0, st, (const float*)dy, (float*)dx, B, Tn, C);
  else if (dt == TS_BF16) ts::launch_k(mean_pool_bwd_kernel<bf16>, grid, 256, 0, st, (const bf16*)dy, (bf16*)dx, B, Tn, C);
  else return set_err(ctx, TS_EDTYPE, "mean_pool: dtype %d", dt);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// MODE 0: y = tanh(x), yd = y * mask      (Dense(activation="tanh") then Dropout, V:1013-1014, V:1046-1047)
// MODE 1: dx = (dy * mask) * (1 - y^2)    (y = the tanh output saved by the forward)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) tanh_drop_kernel(const T* __restrict__ a, const T* __restrict__ y, T* __restrict__ out,
                                                        T* __restrict__ out2, long long n, uint32_t thr, float inv_keep,
                                                        uint64_t seed, const unsigned long long* __restrict__ salt) {
  ts::pdl_enter();
  if (thr) seed = salted_seed(seed, salt);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float mask = thr ? dropout_scale(seed, (uint64_t)i, thr, inv_keep) : 1.f;
    if (MODE == 0) {
      const float t = tanhf(to_f<T>(a[i]));
      const T tr = from_f<T>(t);
      out[i] = tr;
      out2[i] = from_f<T>(to_f<T>(tr) * mask);
    } else {
      const float yy = to_f<T>(y[i]);
      out[i] = from_f<T>(to_f<T>(a[i]) * mask * (1.f - yy * yy));
    }
  }
}

static inline void drop_thr(float drop, uint32_t* thr, float* inv_keep) {
  if (drop <= 0.f) { *thr = 0; *inv_keep = 1.f; return; }
  double t = (double)drop * 4294967296.0;
  if (t > 4294967295.0) t = 4294967295.0;
  *thr = (uint32_t)t;
  *inv_keep = 1.f / (1.f - drop);
}

int tanh_drop_fwd(Ctx* ctx, int dt, const void* x, void* y, void* y_drop, long long n, float drop, uint64_t seed, cudaStream_t st) {
  if (n <= 0) return 0;
  uint32_t thr; float ik;
  drop_thr(drop, &thr, &ik);
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)ctx->num_sms * 16);
  if (dt == TS_F32) ts::launch_k(tanh_drop_kernel<float, 0>, grid, 256, 0, st, (const float*)x, nullptr, (float*)y, (float*)y_drop, n, thr, ik, seed, ctx->d_state);
  else if (dt == TS_BF16) ts::launch_k(tanh_drop_kernel<bf16, 0>, grid, 256, 0, st, (const bf16*)x, nullptr, (bf16*)y, (bf16*)y_drop, n, thr, ik, seed, ctx->d_state);
  else return set_err(ctx, TS_EDTYPE, "tanh_drop: dtype %d", dt);
  TS_LAUNCH_OK(ctx);
  return 0;
}
int tanh_drop_bwd(Ctx* ctx, int dt, const void* dy, const void* y, void* dx, long long n, float drop, uint64_t seed, cudaStream_t st) {
  if (n <= 0) return 0;
  uint32_t thr; float ik;
  drop_thr(drop, &thr, &ik);
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)ctx->num_sms * 16);
  if (dt == TS_F32) ts::launch_k(tanh_drop_kernel<float, 1>, grid, 256, 0, st, (const float*)dy, (const float*)y, (float*)dx, nullptr, n, thr, ik, seed, ctx->d_state);
  else if (dt == TS_BF16) ts::launch_k(tanh_drop_kernel<bf16, 1>, grid, 256, 0, st, (const bf16*)dy, (const bf16*)y, (bf16*)dx, nullptr, n, thr, ik, seed, ctx->d_state);
  else return set_err(ctx, TS_EDTYPE, "tanh_drop: dtype %d", dt);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
