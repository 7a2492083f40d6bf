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
                                                        int left, double* __restrict__ gn_accum, int G) {
  ts::pdl_enter();
  extern __shared__ float sw[];  // wave segment: C0_TILE_T*s + K floats
  __shared__ double sacc[64 * 2]; // GroupNorm statistics of this block's rows (gn_accum != NULL): [G][sum, sum of squares]
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
  if (gn_accum)
    for (int i = threadIdx.x; i < 2 * G; i += 256) sacc[i] = 0.0;
  __syncthreads();
  if (tr < rpi) {
    const int tend = min(C0_TILE_T, T_ - t0);
    float gs = 0.f, gss = 0.f;
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
      // moments of the fp32 values (<= 32 rows x 8 channels per thread; the storage rounding is zero-mean and 2^-9 relative: its
      // effect on a mean over T x C/G >= 10^5 elements is far below the fp32 resolution of the statistics)
#pragma unroll
      for (int i = 0; i < 8; ++i) { gs += o[i]; gss = fmaf(o[i], o[i], gss); }
    }
    if (gn_accum) {
      const int g = (tc * 8) / (C / G);
      atomicAdd(&sacc[2 * g], (double)gs);
      atomicAdd(&sacc[2 * g + 1], (double)gss);
    }
  }
  if (gn_accum) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += 256) atomicAdd(&gn_accum[(long long)b * G * 2 + i], sacc[i]);
  }
}

int conv0_fwd(Ctx* ctx, int dt, const float* wave, const void* w, void* y, long long y_rpb, int B, int N, int T_, int C,
              int k, int s, int left, cudaStream_t st, double* gn_accum, int G) {
  TS_REQUIRE(ctx, !gn_accum || (G > 0 && G <= 64 && C % G == 0 && (C / G) % 8 == 0), TS_ESHAPE, "conv0: GroupNorm statistics need C / G a multiple of 8 (C=%d G=%d)", C, G);
  TS_REQUIRE(ctx, k == 10, TS_EUNSUPPORTED, "conv0: kernel size %d unsupported (reference uses 10)", k);
  TS_REQUIRE(ctx, C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, TS_ESHAPE, "conv0: C=%d unsupported", C);
  dim3 grid(cdiv(T_, C0_TILE_T), B);
  const size_t smem = (C0_TILE_T * s + 10) * sizeof(float);
  if (dt == TS_F32) ts::launch_k(conv0_fwd_kernel<float, 10>, grid, 256, smem, st, wave, (const float*)w, (float*)y, y_rpb, N, T_, C, s, left, gn_accum, G);
  else ts::launch_k(conv0_fwd_kernel<bf16, 10>, grid, 256, smem, st, wave, (const float*)w, (bf16*)y, y_rpb, N, T_, C, s, left, gn_accum, G);
  TS_LAUNCH_OK(ctx);
  return 0;
}

constexpr int C0_WG_ROWS = 512;  // rows reduced per block

// GN = true: dy holds only the GELU-backward product d (pass 1 of the GroupNorm backward) and the gradient of the conv0 output is
// formed on the fly, dz = d * (rstd * gamma) + z * (-rstd^2 * m2) + rstd * (mean * rstd * m2 - m1) with the group means m1, m2 of
// pass 1 (norms.cu gn_gelu_bwd2_kernel's formula) — nothing else consumes dz of layer 0 (no gradient flows into the waveform),
// so the pass that would write it (and this kernel re-reading it) is skipped.
struct Conv0GnPass2 {
  const void* z; long long z_rpb;                 // conv0 output (GroupNorm input), rows per batch
  const float *mean, *rstd, *gamma;
  const double* accum;                            // [B, G, 2] sums of d*gamma and d*gamma*xhat
  int G;
};

constexpr int C0_WG_STAGES = 4;   // cp.async ring depth (two rows of dy [+ z] per thread and stage)


template <typename T, bool GN> struct C0Wg {
  static constexpr int NV = sizeof(T) / 2;                 // 16-byte vectors per 8 elements
  static constexpr int VPS = 2 * (GN ? 2 : 1) * NV;        // vectors per thread and stage
  static size_t smem(int C, int s) { return (size_t)C0_WG_STAGES * VPS * 256 * 16 + (size_t)(C0_WG_ROWS * s + 12 + 10 * C) * sizeof(float); }
};

// The kernel is bound by the dy (and z) bytes it streams, 10 FMAs per loaded element: every thread keeps its own ring of
// cp.async stages in shared memory (it reads back only what it copied itself: no block barriers in the loop), so ~3 stages x 2 rows x
// 16-32 bytes per thread are in flight without holding registers, and two blocks fit an SM (<= 128 registers).
template <typename T, int K, bool GN>
__global__ void __launch_bounds__(256, 2) conv0_wgrad_kernel(const float* __restrict__ wave, const T* __restrict__ dy,
                                                             long long dy_rpb, float* __restrict__ dw, int N, int T_, int C,
                                                             int s, int left, Conv0GnPass2 gn) {
  ts::pdl_enter();
  using W = C0Wg<T, GN>;
  constexpr int NV = W::NV, VPS = W::VPS, S = C0_WG_STAGES, SRC = GN ? 2 : 1;
  extern __shared__ __align__(16) uint8_t smraw[];
  uint4* stg = reinterpret_cast<uint4*>(smraw);                     // [S][VPS][256]
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * C0_WG_ROWS;
  const int seg = C0_WG_ROWS * s + K;
  float* sw = reinterpret_cast<float*>(stg + S * VPS * 256);        // wave segment
  float* sacc = sw + ((seg + 3) & ~3);                              // [K*C] block accumulators
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const int tend = min(C0_WG_ROWS, T_ - t0);
  const T* dyb = dy + ((long long)b * dy_rpb + t0) * C + tc * 8;
  const T* zb = GN ? reinterpret_cast<const T*>(gn.z) + ((long long)b * gn.z_rpb + t0) * C + tc * 8 : nullptr;
  auto issue = [&](int tt, int slot) {
    if (tr < rpi) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int row = tt + u * rpi;
        if (row < tend) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            cp_async16(&stg[((slot * VPS) + (u * SRC + 0) * NV + v) * 256 + threadIdx.x], reinterpret_cast<const uint8_t*>(dyb + (long long)row * C) + 16 * v);
            if (GN) cp_async16(&stg[((slot * VPS) + (u * SRC + 1) * NV + v) * 256 + threadIdx.x], reinterpret_cast<const uint8_t*>(zb + (long long)row * C) + 16 * v);
          }
        }
      }
    }
    cp_async_commit();   // one group per stage, also when empty: the wait below counts groups
  };
#pragma unroll
  for (int p0 = 0; p0 < S - 1; ++p0) issue(tr + p0 * 2 * rpi, p0);   // the ring fills while the wave segment is staged
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
  float ca[8], cb = 0.f, cc = 0.f;
  if (GN && tr < rpi) {
    const int cpg = C / gn.G, g = (tc * 8) / cpg;
    const float mu = gn.mean[b * gn.G + g], rs = gn.rstd[b * gn.G + g];
    const double n = (double)T_ * cpg;
    const float m1 = (float)(gn.accum[((long long)b * gn.G + g) * 2] / n), m2 = (float)(gn.accum[((long long)b * gn.G + g) * 2 + 1] / n);
    load8<float>(gn.gamma + tc * 8, ca);
#pragma unroll
    for (int i = 0; i < 8; ++i) ca[i] *= rs;
    cb = -rs * rs * m2; cc = rs * (mu * rs * m2 - m1);
  }
  int it = 0;
  for (int tt = tr; tt < tend; tt += 2 * rpi, ++it) {
    issue(tt + (S - 1) * 2 * rpi, (it + S - 1) % S);
    cp_async_wait<S - 1>();                    // this iteration's stage has landed (visible to the thread that copied it)
    const int slot = it % S;
    const bool two = tt + rpi < tend;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      float d[8];
      Raw8<T> rd, rz;
      if constexpr (NV == 1) {
        rd.u = stg[((slot * VPS) + (u * SRC + 0)) * 256 + threadIdx.x];
        if (GN) rz.u = stg[((slot * VPS) + (u * SRC + 1)) * 256 + threadIdx.x];
      } else {
        const float4* f = reinterpret_cast<const float4*>(stg);
        rd.a = f[((slot * VPS) + (u * SRC + 0) * NV + 0) * 256 + threadIdx.x]; rd.b = f[((slot * VPS) + (u * SRC + 0) * NV + 1) * 256 + threadIdx.x];
        if (GN) { rz.a = f[((slot * VPS) + (u * SRC + 1) * NV + 0) * 256 + threadIdx.x]; rz.b = f[((slot * VPS) + (u * SRC + 1) * NV + 1) * 256 + threadIdx.x]; }
      }
      unpack8(rd, d);
      if (GN) {
        float z[8];
        unpack8(rz, z);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fmaf(d[i], ca[i], fmaf(z[i], cb, cc));
      }
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float xv = sw[(tt + u * rpi) * s + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(xv, d[i], acc[j][i]);
      }
    }
  }
  cp_async_wait<0>();
  if (tr < rpi) {
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&sacc[j * C + tc * 8 + i], acc[j][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += 256) atomicAdd(&dw[i], sacc[i]);
}

template <typename T, bool GN>
static int conv0_wgrad_launch(Ctx* ctx, const float* wave, const void* dy, long long dy_rpb, float* dw, int B, int N, int T_, int C,
                              int s, int left, const Conv0GnPass2& gn, cudaStream_t st) {
  dim3 grid(cdiv(T_, C0_WG_ROWS), B);
  const size_t smem = C0Wg<T, GN>::smem(C, s);
  TS_REQUIRE(ctx, smem <= 220 * 1024, TS_ESHAPE, "conv0_wgrad: C=%d stride=%d need %zu bytes of shared memory", C, s, smem);
  TS_CUDA_OK(ctx, cudaFuncSetAttribute(conv0_wgrad_kernel<T, 10, GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ts::launch_k(conv0_wgrad_kernel<T, 10, GN>, grid, 256, smem, st, wave, (const T*)dy, dy_rpb, dw, N, T_, C, s, left, gn);
  TS_LAUNCH_OK(ctx);
  return 0;
}

int conv0_wgrad(Ctx* ctx, int dt, const float* wave, const void* dy, long long dy_rpb, float* dw, int B, int N, int T_,
                int C, int k, int s, int left, cudaStream_t st, const void* gn_z, long long gn_z_rpb, const float* gn_mean,
                const float* gn_rstd, const float* gn_gamma, const double* gn_accum, int G) {
  TS_REQUIRE(ctx, k == 10, TS_EUNSUPPORTED, "conv0_wgrad: kernel size %d unsupported", k);
  Conv0GnPass2 gn;
  gn.z = gn_z; gn.z_rpb = gn_z_rpb; gn.mean = gn_mean; gn.rstd = gn_rstd; gn.gamma = gn_gamma; gn.accum = gn_accum; gn.G = G;
  if (gn_z) {
    TS_REQUIRE(ctx, G > 0 && C % G == 0 && (C / G) % 8 == 0, TS_ESHAPE, "conv0_wgrad: fused GroupNorm pass needs C / G a multiple of 8");
    return dt == TS_F32 ? conv0_wgrad_launch<float, true>(ctx, wave, dy, dy_rpb, dw, B, N, T_, C, s, left, gn, st)
                        : conv0_wgrad_launch<bf16, true>(ctx, wave, dy, dy_rpb, dw, B, N, T_, C, s, left, gn, st);
  }
  return dt == TS_F32 ? conv0_wgrad_launch<float, false>(ctx, wave, dy, dy_rpb, dw, B, N, T_, C, s, left, gn, st)
                      : conv0_wgrad_launch<bf16, false>(ctx, wave, dy, dy_rpb, dw, B, N, T_, C, s, left, gn, st);
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
