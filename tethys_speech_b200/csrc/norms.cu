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
  ts::pdl_enter();
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

// backward, part 1: dx (+ dres) — one warp per row, nothing but the row in registers (high occupancy, HBM-bound)
template <typename T, int NCH>
__global__ void __launch_bounds__(256) ln_bwd_dx_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const T* __restrict__ dres,
                                                        T* __restrict__ dx, int rows, int cols) {
  ts::pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float mu = mean[row], rs = rstd[row];
  float xh[NCH][8], d[NCH][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
    if (col < cols) {
      float g[8];
      load8<T>(x + (long long)row * cols + col, xh[ch]);
      load8<T>(dy + (long long)row * cols + col, d[ch]);
      load8<float>(gamma + col, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[ch][i] = (xh[ch][i] - mu) * rs;
        d[ch][i] *= g[i];
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

// backward, part 2: dgamma += sum_rows dy * xhat, dbeta += sum_rows dy. A lane owns 8 consecutive columns, the 8 warps
// of a block interleave the rows of the block's row range; cross-warp reduction in smem, one atomic per column and block.
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_param_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int rows,
                                                           int cols, int rows_per_block) {
  ts::pdl_enter();
  __shared__ float red[2][8][256 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float ag[8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  if (col < cols) {
    for (int r = r0 + w; r < r1; r += 16) {
      float xv[2][8], dv[2][8], mu[2], rs[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rr = r + 8 * u;
        if (rr < r1) {
          load8<T>(x + (long long)rr * cols + col, xv[u]);
          load8<T>(dy + (long long)rr * cols + col, dv[u]);
          mu[u] = mean[rr]; rs[u] = rstd[rr];
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (r + 8 * u >= r1) break;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          ag[i] = fmaf(dv[u][i], (xv[u][i] - mu[u]) * rs[u], ag[i]);
          ab[i] += dv[u][i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[0][w][lane * 8 + i] = ag[i]; red[1][w][lane * 8 + i] = ab[i]; }
  __syncthreads();
  const int c = threadIdx.x;
  float tg = 0.f, tb = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { tg += red[0][i][c]; tb += red[1][i][c]; }
  if (blockIdx.x * 256 + c < cols) {
    atomicAdd(&dgamma[blockIdx.x * 256 + c], tg);
    atomicAdd(&dbeta[blockIdx.x * 256 + c], tb);
  }
}

// backward, fused: ONE pass over (dy, x) produces dx (+ dres), the parameter gradients AND — optionally — the dropped copy
// of dx the next dense layer's backward consumes, with its column sums (= that layer's bias gradient):
//   dx        = rstd * (g o dy - mean(g o dy) - xhat * mean(g o dy o xhat))  [+ dres]
//   dgamma   += sum_rows dy o xhat,  dbeta += sum_rows dy
//   drop_out  = dropout(dx, drop_seed)           (the Dropout in front of the previous Dense in forward order; nullable)
//   drop_csum+= sum_rows drop_out                (nullable)
// A warp owns one row at a time and walks rows block_first + warp, + 8 * gridDim.x, ...; a lane's columns are fixed (8 per
// 256-column chunk), so the per-column partial sums live in registers for the whole kernel and leave through one smem
// reduction over the 8 warps and one atomic per column and block. Replaces ln_bwd_param + ln_bwd_dx (+ dropout + colsum): the
// [rows, cols] tensors are read once instead of two to four times, and 2-4 launches become one.
template <typename T, int NCH, bool DROP>
__global__ void __launch_bounds__(256, 2) ln_bwd_fused_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const T* __restrict__ dres, T* __restrict__ dx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, T* __restrict__ drop_out, float* __restrict__ drop_csum,
                                                           uint32_t thr, float inv_keep, uint64_t seed, const unsigned long long* __restrict__ salt,
                                                           int rows, int cols) {
  ts::pdl_enter();
  __shared__ float red[8][256 + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[NCH][8], ab[NCH][8], ac[NCH][8];   // ac: column sums of the dropped copy (dead code without DROP)
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[ch][i] = 0.f; ab[ch][i] = 0.f; ac[ch][i] = 0.f; }
  }
  DropKey key = flat_drop_key(0, 0);
  if (DROP) key = flat_drop_key(salted_seed(seed, salt), thr);
  for (int row = blockIdx.x * 8 + warp; row < rows; row += 8 * gridDim.x) {
    const float mu = mean[row], rs = rstd[row];
    float xh[NCH][8], d[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int col = ch * 256 + lane * 8;
      if (col < cols) {
        float g[8];                       // gamma stays in L1 (3 KB): re-read per row, 24 registers per thread not pinned
        load8<T>(x + (long long)row * cols + col, xh[ch]);
        load8<T>(dy + (long long)row * cols + col, d[ch]);
        load8<float>(gamma + col, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[ch][i] = (xh[ch][i] - mu) * rs;
          ag[ch][i] = fmaf(d[ch][i], xh[ch][i], ag[ch][i]);
          ab[ch][i] += d[ch][i];
          d[ch][i] *= g[i];
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
        if (DROP) {
          // the dropped copy is taken from the value the consumer would have read: dx rounded to the activation dtype
          round8<T>(o);
          float ds[8];
          dropout_scale8(key, (uint64_t)((long long)row * cols + col), inv_keep, ds);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] *= ds[i];
          store8<T>(drop_out + (long long)row * cols + col, o);
          round8<T>(o);
#pragma unroll
          for (int i = 0; i < 8; ++i) ac[ch][i] += o[i];
        }
      }
    }
  }
  // per-column partials of the 8 warps -> one atomic per column
  auto reduce_to = [&](float (&acc)[NCH][8], float* out) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[ch][i];
      __syncthreads();
      const int c = threadIdx.x;
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][c];
      if (ch * 256 + c < cols) atomicAdd(&out[ch * 256 + c], t);
    }
  };
  reduce_to(ag, dgamma);
  reduce_to(ab, dbeta);
  if (DROP) {
    if (drop_csum) reduce_to(ac, drop_csum);
  }
}

// The same one-pass backward with x / dy / dres streamed through a per-thread cp.async ring (bf16 tensors): S - 1 rows per warp are in
// flight in shared memory while the current one is reduced, instead of one row's loads held in registers (the [6000, 768] passes ran at
// ~2 TB/s, latency-bound). In-place use (dx aliasing dy or dres) stays valid: a row is read and written by the same thread, the read
// of row r is complete (wait_group) before its store, and rows are distinct across warps. The ring's memory is reused for the column
// reduction after the loop.
template <int NCH, bool DROP>
__global__ void __launch_bounds__(256, 2) ln_bwd_ring_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, bf16* __restrict__ drop_out, float* __restrict__ drop_csum,
                                                             uint32_t thr, float inv_keep, uint64_t seed, const unsigned long long* __restrict__ salt,
                                                             int rows, int cols) {
  ts::pdl_enter();
  constexpr int S = 3, VPS = 3 * NCH;                     // per stage: x, dy, dres chunks of one row
  extern __shared__ __align__(16) uint8_t smraw[];
  uint4* stg = reinterpret_cast<uint4*>(smraw);           // [S][VPS][256]
  float (*red)[256 + 8] = reinterpret_cast<float (*)[256 + 8]>(smraw);   // [8][264] after the loop
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[NCH][8], ab[NCH][8], ac[NCH][8];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[ch][i] = 0.f; ab[ch][i] = 0.f; ac[ch][i] = 0.f; }
  }
  DropKey key = flat_drop_key(0, 0);
  if (DROP) key = flat_drop_key(salted_seed(seed, salt), thr);
  const int stride = 8 * gridDim.x;
  auto slot_ptr = [&](int slot, int src, int ch) { return stg + ((slot * VPS) + src * NCH + ch) * 256 + threadIdx.x; };
  auto issue = [&](int row, int slot) {
    const bool rok = row < rows;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int col = ch * 256 + lane * 8;
      const bool ok = rok && col < cols;
      const long long e = ok ? (long long)row * cols + col : 0;
      cp_async16_zfill(slot_ptr(slot, 0, ch), x + e, ok ? 16 : 0);
      cp_async16_zfill(slot_ptr(slot, 1, ch), dy + e, ok ? 16 : 0);
      if (dres) cp_async16_zfill(slot_ptr(slot, 2, ch), dres + e, ok ? 16 : 0);
    }
    cp_async_commit();
  };
  const int row0 = blockIdx.x * 8 + warp;
#pragma unroll
  for (int p0 = 0; p0 < S - 1; ++p0) issue(row0 + p0 * stride, p0);
  int it = 0;
  for (int row = row0; row < rows; row += stride, ++it) {
    issue(row + (S - 1) * stride, (it + S - 1) % S);
    cp_async_wait<S - 1>();
    const int slot = it % S;
    const float mu = mean[row], rs = rstd[row];
    float xh[NCH][8], d[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int col = ch * 256 + lane * 8;
      if (col < cols) {
        float g[8];
        Raw8<bf16> r;
        r.u = *slot_ptr(slot, 0, ch); unpack8(r, xh[ch]);
        r.u = *slot_ptr(slot, 1, ch); unpack8(r, d[ch]);
        load8<float>(gamma + col, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[ch][i] = (xh[ch][i] - mu) * rs;
          ag[ch][i] = fmaf(d[ch][i], xh[ch][i], ag[ch][i]);
          ab[ch][i] += d[ch][i];
          d[ch][i] *= g[i];
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
          float r8[8];
          Raw8<bf16> r;
          r.u = *slot_ptr(slot, 2, ch); unpack8(r, r8);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += r8[i];
        }
        store8<bf16>(dx + (long long)row * cols + col, o);
        if (DROP) {
          round8<bf16>(o);
          float ds[8];
          dropout_scale8(key, (uint64_t)((long long)row * cols + col), inv_keep, ds);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] *= ds[i];
          store8<bf16>(drop_out + (long long)row * cols + col, o);
          round8<bf16>(o);
#pragma unroll
          for (int i = 0; i < 8; ++i) ac[ch][i] += o[i];
        }
      }
    }
  }
  cp_async_wait<0>();
  auto reduce_to = [&](float (&acc)[NCH][8], float* out) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[ch][i];
      __syncthreads();
      const int c = threadIdx.x;
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][c];
      if (ch * 256 + c < cols) atomicAdd(&out[ch * 256 + c], t);
    }
  };
  reduce_to(ag, dgamma);
  reduce_to(ab, dbeta);
  if (DROP) {
    if (drop_csum) reduce_to(ac, drop_csum);
  }
}

template <int N, bool DROP>
static int ln_bwd_ring_launch(Ctx* ctx, int grid, cudaStream_t st, const void* dy, const void* x, const float* gamma, const float* mean,
                              const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, void* drop_out, float* drop_csum,
                              uint32_t thr, float ik, uint64_t seed, int rows, int cols) {
  size_t smem = (size_t)3 * 3 * N * 256 * 16;
  const size_t red = sizeof(float) * 8 * (256 + 8);
  if (smem < red) smem = red;
  TS_CUDA_OK(ctx, cudaFuncSetAttribute(ln_bwd_ring_kernel<N, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ts::launch_k(ln_bwd_ring_kernel<N, DROP>, grid, 256, smem, st, (const bf16*)dy, (const bf16*)x, gamma, mean, rstd, (const bf16*)dres, (bf16*)dx,
               dgamma, dbeta, (bf16*)drop_out, drop_csum, thr, ik, seed, ctx->d_state, rows, cols);
  TS_LAUNCH_OK(ctx);
  return 0;
}

template <typename T>
static int ln_fwd_t(Ctx* ctx, const void* x, const void* res, const float* gamma, const float* beta, void* y,
                    void* sum_out, float* mean, float* rstd, int rows, int cols, float eps, cudaStream_t st) {
  const int nch = cdiv(cols, 256);
  dim3 grid(cdiv(rows, 8));
#define LN_FWD_CASE(N)                                                                                         \
  case N:                                                                                                      \
    ts::launch_k(ln_fwd_kernel<T, N>, grid, 256, 0, st, (const T*)x, (const T*)res, gamma, beta, (T*)y, (T*)sum_out, mean, \
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
                    void* drop_out, float* drop_csum, float drop, uint64_t drop_seed, cudaStream_t st) {
  const int nch = cdiv(cols, 256);
  int grid = cdiv(rows, 8);
  const int cap = ctx->num_sms * 2;      // two resident blocks per SM, every warp walks ~2.5 rows of a [6000, 768] tensor
  if (grid > cap) grid = cap;
  uint32_t thr = 0; float ik = 1.f;
  if (drop_out && drop > 0.f) {
    double t = (double)drop * 4294967296.0;
    thr = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
    ik = 1.f / (1.f - drop);
  }
  // bf16 tensors of a few thousand rows: the cp.async ring version (TETHYS_LN_BWD_DIRECT=1 keeps the register-staged kernel)
  static const bool direct = getenv("TETHYS_LN_BWD_DIRECT") && atoi(getenv("TETHYS_LN_BWD_DIRECT")) != 0;
  if (sizeof(T) == 2 && !direct && nch <= 3 && rows >= 1024) {
#define LN_RING_CASE(N)                                                                                                                    \
  case N:                                                                                                                                  \
    return drop_out ? ln_bwd_ring_launch<N, true>(ctx, grid, st, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, drop_out, drop_csum, thr, \
                                                  ik, drop_seed, rows, cols)                                                               \
                    : ln_bwd_ring_launch<N, false>(ctx, grid, st, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, nullptr, nullptr, 0u, \
                                                   1.f, 0ull, rows, cols);
    switch (nch) { LN_RING_CASE(1) LN_RING_CASE(2) LN_RING_CASE(3) }
#undef LN_RING_CASE
  }
#define LN_BWD_CASE(N)                                                                                                               \
  case N:                                                                                                                            \
    if (drop_out)                                                                                                                    \
      ts::launch_k(ln_bwd_fused_kernel<T, N, true>, grid, 256, 0, st, (const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres, (T*)dx,     \
                                                            dgamma, dbeta, (T*)drop_out, drop_csum, thr, ik, drop_seed, ctx->d_state, \
                                                            rows, cols);                                                             \
    else                                                                                                                             \
      ts::launch_k(ln_bwd_fused_kernel<T, N, false>, grid, 256, 0, st, (const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres, (T*)dx,    \
                                                             dgamma, dbeta, nullptr, nullptr, 0u, 1.f, 0ull, ctx->d_state, rows, cols); \
    break;
  switch (nch) {
    LN_BWD_CASE(1) LN_BWD_CASE(2) LN_BWD_CASE(3) LN_BWD_CASE(4) LN_BWD_CASE(5)
    default: return set_err(ctx, TS_EUNSUPPORTED, "layernorm_bwd: cols=%d > 1280 unsupported", cols);
  }
#undef LN_BWD_CASE
  TS_LAUNCH_OK(ctx);
  return 0;
}

int layernorm_bwd_drop(Ctx* ctx, int dt, const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int cols, void* drop_out, float* drop_csum,
                       float drop, uint64_t drop_seed, cudaStream_t st) {
  TS_REQUIRE(ctx, cols % 8 == 0 && cols > 0 && rows > 0, TS_ESHAPE, "layernorm_bwd: rows=%d cols=%d", rows, cols);
  if (dt == TS_F32) return ln_bwd_t<float>(ctx, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, cols, drop_out, drop_csum, drop, drop_seed, st);
  if (dt == TS_BF16) return ln_bwd_t<bf16>(ctx, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, cols, drop_out, drop_csum, drop, drop_seed, st);
  return set_err(ctx, TS_EDTYPE, "layernorm_bwd: dtype %d", dt);
}

int layernorm_bwd(Ctx* ctx, int dt, const void* dy, const void* x, const float* gamma, const float* mean,
                  const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int cols,
                  cudaStream_t st) {
  return layernorm_bwd_drop(ctx, dt, dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, cols, nullptr, nullptr, 0.f, 0, st);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (+ exact GELU)
// thread <-> 8 consecutive channels of one row; TPR = C/8 threads per row, RPI = 256/TPR rows per block iteration.
// Every thread keeps GN_UNROLL rows (independent 16-byte loads) in flight; the number of rows a block owns is chosen on
// the host so that even the short last conv layers launch several blocks per SM (these kernels are HBM-bound on the
// long layers and latency-bound on the short ones).
// ------------------------------------------------------------------------------------------------
constexpr int GN_UNROLL = 2;

static inline int gn_rows_per_block(Ctx* ctx, long long rows_per_batch, int B, int C) {
  const int rpi = 256 / (C / 8);
  const int step = rpi * GN_UNROLL;
  long long want = (rows_per_batch * B) / ((long long)ctx->num_sms * 6);
  long long rpb = ((want + step - 1) / step) * step;
  if (rpb < step) rpb = step;
  if (rpb > 256 && rpb > step) rpb = 256 >= step ? 256 : step;
  return (int)rpb;
}

__global__ void gn_zero_accum(double* accum, int n) {
  ts::pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) accum[i] = 0.0;
}

template <typename T>
__global__ void __launch_bounds__(256, 4) gn_stats_kernel(const T* __restrict__ x, double* __restrict__ accum, int T_,
                                                       int C, int G, long long rpb, int rows_per_block) {
  ts::pdl_enter();
  __shared__ double sacc[64 * 2];
  const int b = blockIdx.y;
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const int cpg = C / G;
  for (int i = threadIdx.x; i < 2 * G; i += 256) sacc[i] = 0.0;
  __syncthreads();
  float s = 0.f, ss = 0.f;
  const int t0 = blockIdx.x * rows_per_block;
  const int t1 = min(T_, t0 + rows_per_block);
  if (tr < rpi) {
    const T* xb = x + (long long)b * rpb * C + tc * 8;
    for (int t = t0 + tr; t < t1; t += rpi * GN_UNROLL) {
      float v[GN_UNROLL][8];
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u) {
        const int tt = t + u * rpi;
        if (tt < t1) load8<T>(xb + (long long)tt * C, v[u]);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) { s += v[u][i]; ss += v[u][i] * v[u][i]; }
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
  ts::pdl_enter();
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
  ts::launch_k(gn_zero_accum, cdiv(B * G * 2, 256), 256, 0, st, accum, B * G * 2);
  const int rows = gn_rows_per_block(ctx, T_, B, C);
  dim3 grid(cdiv(T_, rows), B);
  if (dt == TS_F32) ts::launch_k(gn_stats_kernel<float>, grid, 256, 0, st, (const float*)x, accum, T_, C, G, rpb, rows);
  else ts::launch_k(gn_stats_kernel<bf16>, grid, 256, 0, st, (const bf16*)x, accum, T_, C, G, rpb, rows);
  ts::launch_k(gn_finalize_kernel, cdiv(B * G, 256), 256, 0, st, accum, mean, rstd, B * G, (double)T_ * (C / G), eps);
  TS_LAUNCH_OK(ctx);
  return 0;
}

template <typename T, bool ACT>   // ACT: GELU after the affine (the feature encoder's pairing, V:248-249); false = GroupNormalization alone
__global__ void __launch_bounds__(256, 4) gn_gelu_fwd_kernel(const T* __restrict__ x, long long x_rpb,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          T* __restrict__ y, long long y_rpb, int y_left, int T_, int C,
                                                          int G, int rows_per_block) {
  ts::pdl_enter();
  const int b = blockIdx.y;
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  if (tr >= rpi) return;
  const int cpg = C / G, g = (tc * 8) / cpg;
  const float mu = mean[b * G + g], rs = rstd[b * G + g];
  float ga[8], be[8];
  load8<float>(gamma + tc * 8, ga);
  load8<float>(beta + tc * 8, be);
  // fold the normalisation into one FMA per element: u = x * a + c
#pragma unroll
  for (int i = 0; i < 8; ++i) { ga[i] *= rs; be[i] = fmaf(-mu, ga[i], be[i]); }
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min((int)y_rpb, r0 + rows_per_block);
  const T* xb = x + (long long)b * x_rpb * C + tc * 8;
  T* yb = y + (long long)b * y_rpb * C + tc * 8;
  // rows of y outside [y_left, y_left + T) are zero padding for the next conv's windows
  const int d0 = max(r0, y_left), d1 = min(r1, y_left + T_);   // data rows of this block
  for (int r = r0 + tr; r < r1; r += rpi) {
    if (r < d0 || r >= d1) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
      store8<T>(yb + r * C, o);
    }
  }
  // main loop: two independent rows in flight per thread, pointer increments only
  int r = d0 + tr;
  const T* xp = xb + (r - y_left) * C;
  T* yp = yb + r * C;
  const int step = rpi * C;
  for (; r + rpi < d1; r += 2 * rpi, xp += 2 * step, yp += 2 * step) {
    float v0[8], v1[8], o[8];
    load8<T>(xp, v0);
    load8<T>(xp + step, v1);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float u = fmaf(v0[i], ga[i], be[i]); o[i] = ACT ? gelu_t<T>(u) : u; }
    store8<T>(yp, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float u = fmaf(v1[i], ga[i], be[i]); o[i] = ACT ? gelu_t<T>(u) : u; }
    store8<T>(yp + step, o);
  }
  if (r < d1) {
    float v0[8], o[8];
    load8<T>(xp, v0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float u = fmaf(v0[i], ga[i], be[i]); o[i] = ACT ? gelu_t<T>(u) : u; }
    store8<T>(yp, o);
  }
}

int groupnorm_gelu_fwd(Ctx* ctx, int dt, const void* x, long long x_rpb, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, void* y, long long y_rpb, int y_left, int B, int T_,
                       int C, int G, cudaStream_t st) {
  const int rows = gn_rows_per_block(ctx, y_rpb, B, C);
  dim3 grid(cdiv(y_rpb, rows), B);
  if (dt == TS_F32)
    ts::launch_k(gn_gelu_fwd_kernel<float, true>, grid, 256, 0, st, (const float*)x, x_rpb, mean, rstd, gamma, beta, (float*)y, y_rpb,
                                                          y_left, T_, C, G, rows);
  else
    ts::launch_k(gn_gelu_fwd_kernel<bf16, true>, grid, 256, 0, st, (const bf16*)x, x_rpb, mean, rstd, gamma, beta, (bf16*)y, y_rpb,
                                                         y_left, T_, C, G, rows);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// GroupNormalization.call alone (V:167-196), no activation: the b-1 sub-layer surface.
int groupnorm_fwd(Ctx* ctx, int dt, const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                  void* y, int B, int T_, int C, int G, cudaStream_t st) {
  const int rows = gn_rows_per_block(ctx, T_, B, C);
  dim3 grid(cdiv(T_, rows), B);
  if (dt == TS_F32)
    ts::launch_k(gn_gelu_fwd_kernel<float, false>, grid, 256, 0, st, (const float*)x, T_, mean, rstd, gamma, beta, (float*)y, T_, 0, T_, C, G, rows);
  else
    ts::launch_k(gn_gelu_fwd_kernel<bf16, false>, grid, 256, 0, st, (const bf16*)x, T_, mean, rstd, gamma, beta, (bf16*)y, T_, 0, T_, C, G, rows);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// Statistics taken by the producer of x (conv0_fwd / the conv GEMM epilogue, ts_gemm_desc.gn_accum) instead of a second pass
// over x: zero the [B, G, 2] fp64 accumulator before the producer, turn it into mean / rstd after it.
int groupnorm_stats_begin(Ctx* ctx, double* accum, int B, int G, cudaStream_t st) {
  ts::launch_k(gn_zero_accum, cdiv(B * G * 2, 256), 256, 0, st, accum, B * G * 2);
  TS_LAUNCH_OK(ctx);
  return 0;
}
int groupnorm_stats_finalize(Ctx* ctx, const double* accum, float* mean, float* rstd, int B, int T_, int C, int G, float eps,
                             cudaStream_t st) {
  ts::launch_k(gn_finalize_kernel, cdiv(B * G, 256), 256, 0, st, accum, mean, rstd, B * G, (double)T_ * (C / G), eps);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// gradient arriving at activation row t: dense `da`, or the col2im gather of the next strided conv's dcol
// (da[b,t,c] = sum_j dcol[b, (t+left-j)/s, j*C + c] over taps j with (t+left-j) % s == 0 and a valid window index).
// KS = k * 8 + s for the specialised (k, s) pairs of the Wav2Vec2 conv stacks; 0 = dense da; -1 = generic.
template <typename T, int KS>
__device__ __forceinline__ void gn_load_upstream(const T* __restrict__ da_row0, const T* __restrict__ dcol_b, const Col2imSrc& col,
                                                 int t, int C, float (&d)[8]) {
  if constexpr (KS == 0) {
    load8<T>(da_row0 + t * C, d);
  } else if constexpr (KS == 3 * 8 + 2) {
    // k = 3, s = 2: an even q = t + left receives taps 0 (window q/2) and 2 (window q/2 - 1), an odd one tap 1 (window (q-1)/2)
    const int q0 = t + col.left, e = q0 & 1, w = q0 >> 1, rowlen = 3 * C;
    float a[8], bb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 0.f; bb[i] = 0.f; }
    if (w < col.t_next) load8<T>(dcol_b + w * rowlen + e * C, a);
    if (e == 0 && w >= 1 && w - 1 < col.t_next) load8<T>(dcol_b + (w - 1) * rowlen + 2 * C, bb);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = a[i] + bb[i];
  } else if constexpr (KS == 2 * 8 + 2) {
    // k = 2, s = 2: exactly one tap (q & 1) of window q / 2
    const int q0 = t + col.left, e = q0 & 1, w = q0 >> 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = 0.f;
    if (w < col.t_next) load8<T>(dcol_b + w * (2 * C) + e * C, d);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = 0.f;
    const long long rowlen = (long long)col.k * C;
    for (int j = 0; j < col.k; ++j) {
      const int q = t + col.left - j;
      if (q < 0) break;
      if (q % col.s) continue;
      const int w = q / col.s;
      if (w >= col.t_next) continue;
      float v[8];
      load8<T>(dcol_b + (long long)w * rowlen + (long long)j * C, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] += v[i];
    }
  }
}

template <typename T, int KS>
__global__ void __launch_bounds__(256, 2) gn_gelu_bwd1_kernel(const T* __restrict__ da, long long da_rpb, Col2imSrc col,
                                                              const T* __restrict__ x, long long x_rpb,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              T* __restrict__ dx, long long dx_rpb, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, double* __restrict__ accum, int T_,
                                                              int C, int G, int rows_per_block) {
  ts::pdl_enter();
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
    const float rs = rstd[b * G + g], nmr = -mean[b * G + g] * rs;
    float ga[8], be[8], ag[8], ab[8];
    load8<float>(gamma + tc * 8, ga);
    load8<float>(beta + tc * 8, be);
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
    const int t0 = blockIdx.x * rows_per_block, t1 = min(T_, t0 + rows_per_block);
    const T* da_b = KS == 0 ? da + (long long)b * da_rpb * C + tc * 8 : nullptr;
    const T* dcol_b = KS == 0 ? nullptr
                              : reinterpret_cast<const T*>(col.dcol) + (long long)b * col.rows_per_batch * ((long long)col.k * C) + tc * 8;
    const T* xb = x + (long long)b * x_rpb * C + tc * 8;
    T* dxb = dx + (long long)b * dx_rpb * C + tc * 8;
    for (int t = t0 + tr; t < t1; t += rpi * GN_UNROLL) {
      float d[GN_UNROLL][8], v[GN_UNROLL][8];
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u) {
        const int tt = t + u * rpi;
        if (tt < t1) {
          gn_load_upstream<T, KS>(da_b, dcol_b, col, tt, C, d[u]);
          load8<T>(xb + tt * C, v[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u) {
        const int tt = t + u * rpi;
        if (tt >= t1) break;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = fmaf(v[u][i], rs, nmr);
          const float dact = d[u][i] * gelu_grad_t<T>(fmaf(ga[i], xh, be[i]));
          o[i] = dact;
          ag[i] = fmaf(dact, xh, ag[i]);
          ab[i] += dact;
        }
        store8<T>(dxb + tt * C, o);
      }
    }
    // group sums for pass 2 follow from the per-channel sums: sum(dact*gamma) and sum(dact*gamma*xhat)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s1 = fmaf(ga[i], ab[i], s1);
      s2 = fmaf(ga[i], ag[i], s2);
      atomicAdd(&sg[tc * 8 + i], ag[i]);
      atomicAdd(&sb[tc * 8 + i], ab[i]);
    }
    atomicAdd(&sacc[2 * g], (double)s1);
    atomicAdd(&sacc[2 * g + 1], (double)s2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) { atomicAdd(&dgamma[i], sg[i]); atomicAdd(&dbeta[i], sb[i]); }
  for (int i = threadIdx.x; i < 2 * G; i += 256) atomicAdd(&accum[(long long)b * G * 2 + i], sacc[i]);
}

// Pass 1 with the upstream gradient / dcol taps and x streamed through a per-thread cp.async ring (KS = 0, 26, 18): the direct
// version above holds every in-flight 16-byte load in registers (122 registers, 2 blocks/SM, 24 % of the warp slots; ncu: 4.4 warps
// stalled on long-scoreboard per issue, 3.3 TB/s). Here up to (S-1) stages x 2 rows x (2-3) vectors per thread are in flight in
// shared memory; out-of-range taps are zero-filled by the copy itself (src-size 0), so the math is the dense path's.
template <typename T, int KS>
__global__ void __launch_bounds__(256, 2) gn_gelu_bwd1_ring_kernel(const T* __restrict__ da, long long da_rpb, Col2imSrc col,
                                                                   const T* __restrict__ x, long long x_rpb,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   T* __restrict__ dx, long long dx_rpb, float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta, double* __restrict__ accum, int T_,
                                                                   int C, int G, int rows_per_block) {
  ts::pdl_enter();
  static_assert(KS == 0 || KS == 26 || KS == 18, "ring version: dense upstream, (k 3, s 2) or (k 2, s 2) taps");
  constexpr int NV = sizeof(T) / 2;               // 16-byte vectors per 8 elements
  constexpr int NSRC = KS == 26 ? 3 : 2;          // tap a, [tap b,] x
  constexpr int VPS = 2 * NSRC * NV;              // vectors per thread and stage (two rows)
  constexpr int S = sizeof(T) == 2 ? 4 : 2;       // ring depth
  extern __shared__ __align__(16) uint8_t smraw[];
  uint4* stg = reinterpret_cast<uint4*>(smraw);                      // [S][VPS][256]
  float* sg = reinterpret_cast<float*>(stg + S * VPS * 256);         // [C] dgamma partials
  float* sb = sg + C;                                                // [C] dbeta partials
  double* sacc = reinterpret_cast<double*>(sb + C);                  // [2 G]
  const int b = blockIdx.y;
  const int tpr = C / 8, rpi = 256 / tpr;
  const int tc = threadIdx.x % tpr, tr = threadIdx.x / tpr;
  const int cpg = C / G, g = (tc * 8) / cpg;
  const int t0 = blockIdx.x * rows_per_block, t1 = min(T_, t0 + rows_per_block);
  const bool active = tr < rpi;
  const T* da_b = KS == 0 ? da + (long long)b * da_rpb * C + tc * 8 : nullptr;
  const T* dcol_b = KS == 0 ? nullptr
                            : reinterpret_cast<const T*>(col.dcol) + (long long)b * col.rows_per_batch * ((long long)col.k * C) + tc * 8;
  const T* xb = x + (long long)b * x_rpb * C + tc * 8;
  auto slot_ptr = [&](int slot, int u, int src) { return stg + ((slot * VPS) + (u * NSRC + src) * NV) * 256 + threadIdx.x; };
  auto copy8 = [&](uint4* dst, const T* src, bool ok) {
#pragma unroll
    for (int v = 0; v < NV; ++v) cp_async16_zfill(dst + v * 256, reinterpret_cast<const uint8_t*>(ok ? src : xb) + (ok ? 16 * v : 0), ok ? 16 : 0);
  };
  auto issue = [&](int t, int slot) {
    if (active) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int tt = t + u * rpi;
        const bool inb = tt < t1;
        if constexpr (KS == 0) {
          copy8(slot_ptr(slot, u, 0), da_b + (long long)tt * C, inb);
        } else if constexpr (KS == 26) {
          // k = 3, s = 2: an even q = t + left receives taps 0 (window q/2) and 2 (window q/2 - 1), an odd one tap 1 (window (q-1)/2)
          const int q0 = tt + col.left, e = q0 & 1, w = q0 >> 1;
          const long long rowlen = 3ll * C;
          copy8(slot_ptr(slot, u, 0), dcol_b + w * rowlen + e * C, inb && w < col.t_next);
          copy8(slot_ptr(slot, u, 1), dcol_b + (w - 1) * rowlen + 2 * C, inb && e == 0 && w >= 1 && w - 1 < col.t_next);
        } else {
          // k = 2, s = 2: exactly one tap (q & 1) of window q / 2
          const int q0 = tt + col.left, e = q0 & 1, w = q0 >> 1;
          copy8(slot_ptr(slot, u, 0), dcol_b + (long long)w * (2 * C) + e * C, inb && w < col.t_next);
        }
        copy8(slot_ptr(slot, u, NSRC - 1), xb + (long long)tt * C, inb);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int p0 = 0; p0 < S - 1; ++p0) issue(t0 + tr + p0 * 2 * rpi, p0);
  for (int i = threadIdx.x; i < 2 * C; i += 256) sg[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * G; i += 256) sacc[i] = 0.0;
  __syncthreads();
  float ga[8], be[8], ag[8], ab[8];
  float rs = 0.f, nmr = 0.f;
  if (active) {
    rs = rstd[b * G + g]; nmr = -mean[b * G + g] * rs;
    load8<float>(gamma + tc * 8, ga);
    load8<float>(beta + tc * 8, be);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  T* dxb = dx + (long long)b * dx_rpb * C + tc * 8;
  int it = 0;
  for (int t = t0 + tr; t < t1; t += 2 * rpi, ++it) {
    issue(t + (S - 1) * 2 * rpi, (it + S - 1) % S);
    cp_async_wait<S - 1>();
    const int slot = it % S;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int tt = t + u * rpi;
      if (tt >= t1) break;
      float d[8], v[8], o[8];
      Raw8<T> r;
      load_raw8_smem(slot_ptr(slot, u, 0), 256, r);
      unpack8(r, d);
      if constexpr (KS == 26) {
        float d2[8];
        load_raw8_smem(slot_ptr(slot, u, 1), 256, r);
        unpack8(r, d2);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] += d2[i];
      }
      load_raw8_smem(slot_ptr(slot, u, NSRC - 1), 256, r);
      unpack8(r, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = fmaf(v[i], rs, nmr);
        const float dact = d[i] * gelu_grad_t<T>(fmaf(ga[i], xh, be[i]));
        o[i] = dact;
        ag[i] = fmaf(dact, xh, ag[i]);
        ab[i] += dact;
      }
      store8<T>(dxb + (long long)tt * C, o);
    }
  }
  cp_async_wait<0>();
  if (active) {
    // group sums for pass 2 follow from the per-channel sums: sum(dact*gamma) and sum(dact*gamma*xhat)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s1 = fmaf(ga[i], ab[i], s1);
      s2 = fmaf(ga[i], ag[i], s2);
      atomicAdd(&sg[tc * 8 + i], ag[i]);
      atomicAdd(&sb[tc * 8 + i], ab[i]);
    }
    atomicAdd(&sacc[2 * g], (double)s1);
    atomicAdd(&sacc[2 * g + 1], (double)s2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) { atomicAdd(&dgamma[i], sg[i]); atomicAdd(&dbeta[i], sb[i]); }
  for (int i = threadIdx.x; i < 2 * G; i += 256) atomicAdd(&accum[(long long)b * G * 2 + i], sacc[i]);
}

template <typename T>
__global__ void __launch_bounds__(256, 4) gn_gelu_bwd2_kernel(const T* __restrict__ x, long long x_rpb,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, T* __restrict__ dx,
                                                           long long dx_rpb, const double* __restrict__ accum, int T_, int C,
                                                           int G, int rows_per_block) {
  ts::pdl_enter();
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
  const int r0 = blockIdx.x * rows_per_block, r1 = min((int)dx_rpb, r0 + rows_per_block);
  const T* xb = x + (long long)b * x_rpb * C + tc * 8;
  T* db = dx + (long long)b * dx_rpb * C + tc * 8;
  const int d1 = min(r1, T_);
  for (int r = max(r0, T_) + tr; r < r1; r += rpi) {  // rows beyond T are zeroed (window slack of the conv GEMMs)
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    store8<T>(db + r * C, o);
  }
  // dx = rs * (d * gamma - m1 - xhat * m2) = d * (rs * gamma) + x * (-rs * rs * m2) + (rs * (mu * rs * m2 - m1))
  float ca[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ca[i] = rs * ga[i];
  const float cb = -rs * rs * m2, cc = rs * (mu * rs * m2 - m1);
  int r = r0 + tr;
  const T* xp = xb + r * C;
  T* dp = db + r * C;
  const int step = rpi * C;
  for (; r + rpi < d1; r += 2 * rpi, xp += 2 * step, dp += 2 * step) {
    float v0[8], v1[8], g0[8], g1[8], o[8];
    load8<T>(xp, v0); load8<T>(dp, g0);
    load8<T>(xp + step, v1); load8<T>(dp + step, g1);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(g0[i], ca[i], fmaf(v0[i], cb, cc));
    store8<T>(dp, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(g1[i], ca[i], fmaf(v1[i], cb, cc));
    store8<T>(dp + step, o);
  }
  if (r < d1) {
    float v0[8], g0[8], o[8];
    load8<T>(xp, v0); load8<T>(dp, g0);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(g0[i], ca[i], fmaf(v0[i], cb, cc));
    store8<T>(dp, o);
  }
}

// ring version (cp.async staging) for the specialised tap patterns; TETHYS_GN_BWD_DIRECT=1 keeps the register-staged kernel
template <typename TT, int KS>
static int gn_bwd1_launch(Ctx* ctx, dim3 g1, size_t smem, cudaStream_t st, const void* da, long long da_rpb, const Col2imSrc& c, const void* x,
                          long long x_rpb, const float* mean, const float* rstd, const float* gamma, const float* beta, void* dx,
                          long long dx_rpb, float* dgamma, float* dbeta, double* accum, int T_, int C, int G, int rows1) {
  static const bool direct = getenv("TETHYS_GN_BWD_DIRECT") && atoi(getenv("TETHYS_GN_BWD_DIRECT")) != 0;
  if constexpr (KS == 0 || KS == 26 || KS == 18) {
    constexpr int NSRC_ = KS == 26 ? 3 : 2, S_ = sizeof(TT) == 2 ? 4 : 2;
    const size_t rsmem = (size_t)S_ * 2 * NSRC_ * (sizeof(TT) / 2) * 256 * 16 + smem;
    if (!direct && rsmem <= 200 * 1024) {
      TS_CUDA_OK(ctx, cudaFuncSetAttribute(gn_gelu_bwd1_ring_kernel<TT, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
      ts::launch_k(gn_gelu_bwd1_ring_kernel<TT, KS>, g1, 256, rsmem, st, (const TT*)da, da_rpb, c, (const TT*)x, x_rpb, mean, rstd, gamma, beta,
                   (TT*)dx, dx_rpb, dgamma, dbeta, accum, T_, C, G, rows1);
      return 0;
    }
  }
  ts::launch_k(gn_gelu_bwd1_kernel<TT, KS>, g1, 256, smem, st, (const TT*)da, da_rpb, c, (const TT*)x, x_rpb, mean, rstd, gamma, beta, (TT*)dx,
               dx_rpb, dgamma, dbeta, accum, T_, C, G, rows1);
  return 0;
}
#define TS_TRY_RC(...) do { int _rc = (__VA_ARGS__); if (_rc) return _rc; } while (0)

int groupnorm_gelu_bwd(Ctx* ctx, int dt, const void* da, long long da_rpb, const Col2imSrc* col, const void* x,
                       long long x_rpb, const float* mean, const float* rstd, const float* gamma, const float* beta,
                       void* dx, long long dx_rpb, float* dgamma, float* dbeta, double* accum, int B, int T_, int C, int G,
                       cudaStream_t st, bool skip_pass2) {
  Col2imSrc c;
  if (col) c = *col; else { c.dcol = nullptr; c.rows_per_batch = 0; c.t_next = 0; c.k = 0; c.s = 1; c.left = 0; }
  ts::launch_k(gn_zero_accum, cdiv(B * G * 2, 256), 256, 0, st, accum, B * G * 2);
  const size_t smem = 2 * C * sizeof(float) + 2 * G * sizeof(double);
  const int rows1 = gn_rows_per_block(ctx, T_, B, C), rows2 = gn_rows_per_block(ctx, dx_rpb, B, C);
  dim3 g1(cdiv(T_, rows1), B), g2(cdiv(dx_rpb, rows2), B);
  int ks = -1;
  if (!c.dcol) ks = 0;
  else if (c.k == 3 && c.s == 2) ks = 3 * 8 + 2;
  else if (c.k == 2 && c.s == 2) ks = 2 * 8 + 2;
#define GN_BWD1(TT, KS) TS_TRY_RC(gn_bwd1_launch<TT, KS>(ctx, g1, smem, st, da, da_rpb, c, x, x_rpb, mean, rstd, gamma, beta, dx, dx_rpb, dgamma, dbeta, \
                                                         accum, T_, C, G, rows1))
  if (dt == TS_F32) {
    if (ks == 0) GN_BWD1(float, 0); else if (ks == 26) GN_BWD1(float, 26); else if (ks == 18) GN_BWD1(float, 18); else GN_BWD1(float, -1);
    if (!skip_pass2) ts::launch_k(gn_gelu_bwd2_kernel<float>, g2, 256, 0, st, (const float*)x, x_rpb, mean, rstd, gamma, (float*)dx, dx_rpb, accum,
                                                   T_, C, G, rows2);
  } else {
    if (ks == 0) GN_BWD1(bf16, 0); else if (ks == 26) GN_BWD1(bf16, 26); else if (ks == 18) GN_BWD1(bf16, 18); else GN_BWD1(bf16, -1);
    if (!skip_pass2) ts::launch_k(gn_gelu_bwd2_kernel<bf16>, g2, 256, 0, st, (const bf16*)x, x_rpb, mean, rstd, gamma, (bf16*)dx, dx_rpb, accum, T_,
                                                  C, G, rows2);
  }
#undef GN_BWD1
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
