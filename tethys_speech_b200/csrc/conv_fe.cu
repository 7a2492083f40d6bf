// K4 Wav2Vec2 conv0 (Cin = 1, k = 10, s = 5, TF SAME padding; V:240-247) forward + weight gradient on CUDA cores
// (the op is bound by the 512-channel fan-out written to HBM, not by math), and the group-major repacks that turn
// the grouped positional conv (K7, V:271-277: k = 128, 16 groups of 32 -> 32 channels, SAME 63/64) into
// window-GEMMs for the tcgen05 engine (rows of the A operand overlap: lda = C/G, K = k*C/G).
#include "ops.cuh"
#include "vec.cuh"

namespace ts {

constexpr int C0_TILE_T = 128;  // output time steps per block

template <typename T, int K>
__global__ void __launch_bounds__(256) conv0_fwd_kernel(const float* __restrict__ wave, const float* __restrict__ w,
                                                        T* __restrict__ y, long long y_rpb, int N, int T_, int C, int s,
                                                        int left) {
  ts::pdl_enter();
  extern __shared__ float sw[];  // wave segment: C0_TILE_T*s + K floats
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * C0_TILE_T;
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const int seg = C0_TILE_T * s + K;
  const long long base = (long long)t0 * s - left;
  for (int i = threadIdx.x; i < seg; i += 256) {
    const long long n = base + i;
    sw[i] = (n >= 0 && n < N) ? wave[(long long)b * N + n] : 0.f;
  }
  float wr[K][8];
#pragma unroll
  for (int j = 0; j < K; ++j) load8<float>(w + (long long)j * C + tc * 8, wr[j]);
  __syncthreads();
  if (tr >= rpi) return;
  const int tend = min(C0_TILE_T, T_ - t0);
  for (int tt = tr; tt < tend; tt += rpi) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float xv = sw[tt * s + j];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(xv, wr[j][i], o[i]);
    }
    store8<T>(y + ((long long)b * y_rpb + t0 + tt) * C + tc * 8, o);
  }
}

int conv0_fwd(Ctx* ctx, int dt, const float* wave, const void* w, void* y, long long y_rpb, int B, int N, int T_, int C,
              int k, int s, int left, cudaStream_t st) {
  TS_REQUIRE(ctx, k == 10, TS_EUNSUPPORTED, "conv0: kernel size %d unsupported (reference uses 10)", k);
  TS_REQUIRE(ctx, C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, TS_ESHAPE, "conv0: C=%d unsupported", C);
  dim3 grid(cdiv(T_, C0_TILE_T), B);
  const size_t smem = (C0_TILE_T * s + 10) * sizeof(float);
  if (dt == TS_F32) ts::launch_k(conv0_fwd_kernel<float, 10>, grid, 256, smem, st, wave, (const float*)w, (float*)y, y_rpb, N, T_, C, s, left);
  else ts::launch_k(conv0_fwd_kernel<bf16, 10>, grid, 256, smem, st, wave, (const float*)w, (bf16*)y, y_rpb, N, T_, C, s, left);
  TS_LAUNCH_OK(ctx);
  return 0;
}

constexpr int C0_WG_ROWS = 512;  // rows reduced per block

template <typename T, int K>
__global__ void __launch_bounds__(256) conv0_wgrad_kernel(const float* __restrict__ wave, const T* __restrict__ dy,
                                                          long long dy_rpb, float* __restrict__ dw, int N, int T_, int C,
                                                          int s, int left) {
  ts::pdl_enter();
  extern __shared__ float sm[];  // wave segment [C0_WG_ROWS*s + K] then accumulators [K*C]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * C0_WG_ROWS;
  const int seg = C0_WG_ROWS * s + K;
  float* sw = sm;
  float* sacc = sm + seg;
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const long long base = (long long)t0 * s - left;
  for (int i = threadIdx.x; i < seg; i += 256) {
    const long long n = base + i;
    sw[i] = (n >= 0 && n < N) ? wave[(long long)b * N + n] : 0.f;
  }
  for (int i = threadIdx.x; i < K * C; i += 256) sacc[i] = 0.f;
  __syncthreads();
  float acc[K][8];
#pragma unroll
  for (int j = 0; j < K; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
  const int tend = min(C0_WG_ROWS, T_ - t0);
  if (tr < rpi) {
    const T* dyb = dy + ((long long)b * dy_rpb + t0) * C + tc * 8;
    for (int tt = tr; tt < tend; tt += 2 * rpi) {  // two independent 16-byte loads in flight per thread
      float d[2][8];
      const bool two = tt + rpi < tend;
      load8<T>(dyb + (long long)tt * C, d[0]);
      if (two) load8<T>(dyb + (long long)(tt + rpi) * C, d[1]);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float xv = sw[(tt + u * rpi) * s + j];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(xv, d[u][i], acc[j][i]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&sacc[j * C + tc * 8 + i], acc[j][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += 256) atomicAdd(&dw[i], sacc[i]);
}

int conv0_wgrad(Ctx* ctx, int dt, const float* wave, const void* dy, long long dy_rpb, float* dw, int B, int N, int T_,
                int C, int k, int s, int left, cudaStream_t st) {
  TS_REQUIRE(ctx, k == 10, TS_EUNSUPPORTED, "conv0_wgrad: kernel size %d unsupported", k);
  dim3 grid(cdiv(T_, C0_WG_ROWS), B);
  const size_t smem = (C0_WG_ROWS * s + 10 + 10 * C) * sizeof(float);
  if (dt == TS_F32) {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(conv0_wgrad_kernel<float, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); set = true; }
    ts::launch_k(conv0_wgrad_kernel<float, 10>, grid, 256, smem, st, wave, (const float*)dy, dy_rpb, dw, N, T_, C, s, left);
  } else {
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(conv0_wgrad_kernel<bf16, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); set = true; }
    ts::launch_k(conv0_wgrad_kernel<bf16, 10>, grid, 256, smem, st, wave, (const bf16*)dy, dy_rpb, dw, N, T_, C, s, left);
  }
  TS_LAUNCH_OK(ctx);
  return 0;
}

// x [B,T,C] -> xg [G][B][R = T+K-1][cpg]; rows [left, left+T) carry data, the rest are zero.
template <typename T>
__global__ void __launch_bounds__(256) posconv_pack_kernel(const T* __restrict__ x, T* __restrict__ xg, int B, int T_,
                                                           int C, int G, int R, int left) {
  ts::pdl_enter();
  const int cpg = C / G, v8 = cpg / 8;
  const long long total = (long long)G * B * R * v8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % v8);
    long long r_ = i / v8;
    const int r = (int)(r_ % R);
    r_ /= R;
    const int b = (int)(r_ % B), g = (int)(r_ / B);
    const int t = r - left;
    float v[8];
    if (t >= 0 && t < T_) load8<T>(x + ((long long)b * T_ + t) * C + g * cpg + c8 * 8, v);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    store8<T>(xg + i * 8, v);
  }
}
int posconv_pack(Ctx* ctx, int dt, const void* x, void* xg, int B, int T_, int C, int G, int K, int left, cudaStream_t st) {
  TS_REQUIRE(ctx, C % G == 0 && (C / G) % 8 == 0, TS_ESHAPE, "posconv_pack: C=%d G=%d", C, G);
  const int R = T_ + K - 1;
  const long long total = (long long)G * B * R * (C / G / 8);
  const int grid = (int)min((total + 255) / 256, (long long)ctx->num_sms * 16);
  if (dt == TS_F32) ts::launch_k(posconv_pack_kernel<float>, grid, 256, 0, st, (const float*)x, (float*)xg, B, T_, C, G, R, left);
  else ts::launch_k(posconv_pack_kernel<bf16>, grid, 256, 0, st, (const bf16*)x, (bf16*)xg, B, T_, C, G, R, left);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// wt[g][K-1-j][o][c] = w[j][c][g*cpg+o]
template <typename T>
__global__ void posconv_flip_kernel(const T* __restrict__ w, T* __restrict__ wt, int K, int C, int G) {
  ts::pdl_enter();
  const int cpg = C / G;
  const long long total = (long long)K * cpg * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes wt: (((g*K + jj)*cpg + o)*cpg + c)
    const int c = (int)(i % cpg);
    long long r = i / cpg;
    const int o = (int)(r % cpg);
    r /= cpg;
    const int jj = (int)(r % K), g = (int)(r / K);
    const int j = K - 1 - jj;
    wt[i] = w[((long long)j * cpg + c) * C + g * cpg + o];
  }
}
int posconv_flip_weight(Ctx* ctx, int dt, const void* w, void* wt, int K, int C, int G, cudaStream_t st) {
  const long long total = (long long)K * (C / G) * C;
  const int grid = (int)min((total + 255) / 256, (long long)ctx->num_sms * 16);
  if (dt == TS_F32) ts::launch_k(posconv_flip_kernel<float>, grid, 256, 0, st, (const float*)w, (float*)wt, K, C, G);
  else ts::launch_k(posconv_flip_kernel<bf16>, grid, 256, 0, st, (const bf16*)w, (bf16*)wt, K, C, G);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
