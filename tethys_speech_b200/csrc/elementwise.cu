// K12 GELU fwd/bwd, dropout, casts, bias-gradient column sums, transposes: pure HBM-bound helpers
// (W:195,203-205,329,333,336,342; V:132-136,281,393-396,431). Grid-stride, 16-byte vectors.
#include "ops.cuh"
#include "vec.cuh"

namespace ts {

static inline int ew_grid(Ctx* ctx, long long nvec) {
  long long g = (nvec + 255) / 256;
  const long long cap = (long long)ctx->num_sms * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__global__ void cast_kernel(const float* __restrict__ s, bf16* __restrict__ d, long long n) {
  ts::pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float v[8];
      load8<float>(s + i, v);
      store8<bf16>(d + i, v);
    } else {
      for (long long j = i; j < n; ++j) d[j] = __float2bfloat16_rn(s[j]);
    }
  }
}
int cast_f32_to_bf16(Ctx* ctx, const float* src, void* dst, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  ts::launch_k(cast_kernel, ew_grid(ctx, (n + 7) / 8), 256, 0, st, src, (bf16*)dst, n);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// gradient buckets for the all-reduce in bf16 ("perf mode" of SURVEY §8e): dst = bf16(src * scale[0]) and back
__global__ void __launch_bounds__(256) grad_pack_kernel(const float* __restrict__ s, bf16* __restrict__ d, long long n,
                                                        const float* __restrict__ scale) {
  ts::pdl_enter();
  const float sc = scale ? __ldg(scale) : 1.f;
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float v[8];
      load8<float>(s + i, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= sc;
      store8<bf16>(d + i, v);
    } else {
      for (long long j = i; j < n; ++j) d[j] = __float2bfloat16_rn(s[j] * sc);
    }
  }
}
__global__ void __launch_bounds__(256) grad_unpack_kernel(const bf16* __restrict__ s, float* __restrict__ d, long long n) {
  ts::pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float v[8];
      load8<bf16>(s + i, v);
      store8<float>(d + i, v);
    } else {
      for (long long j = i; j < n; ++j) d[j] = __bfloat162float(s[j]);
    }
  }
}
int grad_pack_bf16(Ctx* ctx, const float* src, void* dst, long long n, const float* scale_dev, cudaStream_t st) {
  if (n <= 0) return 0;
  TS_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(src) & 31) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0), TS_EINVAL,
             "grad_pack_bf16: buckets must start on a 32-byte (fp32) / 16-byte (bf16) boundary");
  ts::launch_k(grad_pack_kernel, ew_grid(ctx, (n + 7) / 8), 256, 0, st, src, (bf16*)dst, n, scale_dev);
  TS_LAUNCH_OK(ctx);
  return 0;
}
int grad_unpack_bf16(Ctx* ctx, const void* src, float* dst, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  TS_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0), TS_EINVAL,
             "grad_unpack_bf16: buckets must start on a 32-byte (fp32) / 16-byte (bf16) boundary");
  ts::launch_k(grad_unpack_kernel, ew_grid(ctx, (n + 7) / 8), 256, 0, st, (const bf16*)src, dst, n);
  TS_LAUNCH_OK(ctx);
  return 0;
}

int fill_zero(Ctx* ctx, void* p, long long bytes, cudaStream_t st) {
  if (bytes <= 0) return 0;
  TS_CUDA_OK(ctx, cudaMemsetAsync(p, 0, (size_t)bytes, st));
  return 0;
}

// mode 0: out = gelu(a)*mask ; mode 1: out = a*gelu'(b)*mask ; mode 2: out = a*mask ; mode 3: out = a+b
template <typename T, int MODE>
__global__ void __launch_bounds__(256) ew_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                 long long n, uint32_t thr, float inv_keep, uint64_t seed,
                                                 const unsigned long long* __restrict__ salt) {
  ts::pdl_enter();
  if (MODE != 3 && thr) seed = salted_seed(seed, salt);
  const DropKey key = flat_drop_key(seed, thr);
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    float va[8], vb[8], o[8];
    if (i + 8 <= n) {
      load8<T>(a + i, va);
      if (MODE == 1 || MODE == 3) load8<T>(b + i, vb);
    } else {
      for (int j = 0; j < 8; ++j) {
        va[j] = (i + j < n) ? to_f<T>(a[i + j]) : 0.f;
        vb[j] = ((MODE == 1 || MODE == 3) && i + j < n) ? to_f<T>(b[i + j]) : 0.f;
      }
    }
    float ds[8];
    if (MODE != 3 && thr) dropout_scale8(key, (uint64_t)i, inv_keep, ds);   // i is a multiple of 8: one chunk hash for the 8 elements
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float r;
      if (MODE == 0) r = gelu_t<T>(va[j]);
      else if (MODE == 1) r = va[j] * gelu_grad_t<T>(vb[j]);
      else if (MODE == 2) r = va[j];
      else r = va[j] + vb[j];
      if (MODE != 3 && thr) r *= ds[j];
      o[j] = r;
    }
    if (i + 8 <= n) store8<T>(out + i, o);
    else
      for (int j = 0; j < 8 && i + j < n; ++j) out[i + j] = from_f<T>(o[j]);
  }
}

static inline void drop_params(float drop, uint32_t* thr, float* inv_keep) {
  if (drop <= 0.f) { *thr = 0; *inv_keep = 1.f; return; }
  double t = (double)drop * 4294967296.0;
  if (t > 4294967295.0) t = 4294967295.0;
  *thr = (uint32_t)t;
  *inv_keep = 1.f / (1.f - drop);
}

template <int MODE>
static int ew_launch(Ctx* ctx, int dt, const void* a, const void* b, void* out, long long n, float drop, uint64_t seed,
                     cudaStream_t st) {
  if (n <= 0) return 0;
  uint32_t thr; float ik;
  drop_params(drop, &thr, &ik);
  const int grid = ew_grid(ctx, (n + 7) / 8);
  if (dt == TS_F32) ts::launch_k(ew_kernel<float, MODE>, grid, 256, 0, st, (const float*)a, (const float*)b, (float*)out, n, thr, ik, seed, ctx->d_state);
  else if (dt == TS_BF16) ts::launch_k(ew_kernel<bf16, MODE>, grid, 256, 0, st, (const bf16*)a, (const bf16*)b, (bf16*)out, n, thr, ik, seed, ctx->d_state);
  else return set_err(ctx, TS_EDTYPE, "elementwise: dtype %d", dt);
  TS_LAUNCH_OK(ctx);
  return 0;
}

int gelu_fwd(Ctx* ctx, int dt, const void* u, void* out, long long n, float drop, uint64_t seed, cudaStream_t st) {
  return ew_launch<0>(ctx, dt, u, nullptr, out, n, drop, seed, st);
}
int gelu_bwd(Ctx* ctx, int dt, const void* df, const void* u, void* du, long long n, float drop, uint64_t seed, cudaStream_t st) {
  return ew_launch<1>(ctx, dt, df, u, du, n, drop, seed, st);
}
int dropout_apply(Ctx* ctx, int dt, const void* x, void* y, long long n, float drop, uint64_t seed, cudaStream_t st) {
  return ew_launch<2>(ctx, dt, x, nullptr, y, n, drop, seed, st);
}
int add_tensors(Ctx* ctx, int dt, const void* a, const void* b, void* y, long long n, cudaStream_t st) {
  return ew_launch<3>(ctx, dt, a, b, y, n, 0.f, 0, st);
}

// column sums: block = 32 x 8 threads over a [rows-chunk, 32-col] slab; smem transpose-free: each thread
// strides rows, warp lanes map to consecutive columns (coalesced), cross-warp reduce in smem, one atomic per col.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ld, int rows, int cols,
                                                     float* __restrict__ out, int rows_per_block) {
  ts::pdl_enter();
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (col < cols)
    for (int r = r0 + w; r < r1; r += 8) s += to_f<T>(x[(long long)r * ld + col]);
  red[w][lane] = s;
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    if (col < cols) atomicAdd(&out[col], t);
  }
}
// vectorised variant (cols % 8 == 0, 16-byte aligned rows): a lane owns 8 consecutive columns (one 16-byte load per
// row), a warp a 256-column strip, the 8 warps of a block interleave rows; 4 independent loads in flight per thread.
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, long long ld, int rows, int cols,
                                                         float* __restrict__ out, int rows_per_block) {
  ts::pdl_enter();
  __shared__ float red[8][256 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < cols) {
    int r = r0 + w;
    for (; r + 24 < r1; r += 32) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8<T>(x + (long long)(r + 8 * u) * ld + col, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v[u][i];
    }
    for (; r < r1; r += 8) {
      float v[8];
      load8<T>(x + (long long)r * ld + col, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i][c];
  if (blockIdx.x * 256 + c < cols) atomicAdd(&out[blockIdx.x * 256 + c], t);
}
int colsum_acc(Ctx* ctx, int dt, const void* x, long long ld, int rows, int cols, float* out, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  const int esz = dt == TS_F32 ? 4 : 2;
  if (cols % 8 == 0 && (ld * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int cb = cdiv(cols, 256);
    int rb = (ctx->num_sms * 3) / cb;
    if (rb < 1) rb = 1;
    int rpb = cdiv(rows, rb);
    rpb = ((rpb + 31) / 32) * 32;
    dim3 grid(cb, cdiv(rows, rpb));
    if (dt == TS_F32) ts::launch_k(colsum_vec_kernel<float>, grid, 256, 0, st, (const float*)x, ld, rows, cols, out, rpb);
    else ts::launch_k(colsum_vec_kernel<bf16>, grid, 256, 0, st, (const bf16*)x, ld, rows, cols, out, rpb);
    TS_LAUNCH_OK(ctx);
    return 0;
  }
  const int rpb = 512;
  dim3 grid(cdiv(cols, 32), cdiv(rows, rpb));
  if (dt == TS_F32) ts::launch_k(colsum_kernel<float>, grid, 256, 0, st, (const float*)x, ld, rows, cols, out, rpb);
  else ts::launch_k(colsum_kernel<bf16>, grid, 256, 0, st, (const bf16*)x, ld, rows, cols, out, rpb);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// Element-wise pass fused with the column sums of its OUTPUT (= the bias gradient of the Dense layer whose dY this output is):
//   mode 1: out = a * gelu'(b) * mask   (backward of GELU + dropout: dU of fc1 from dF, W:194-205 / V:391-396)
//   mode 2: out = a * mask              (backward of a Dropout: the dY of the Dense in front of it)
// Dense [rows, cols] tensors, cols % 8 == 0. Same thread layout as colsum_vec_kernel (a lane owns 8 consecutive columns, the 8
// warps of a block interleave the rows of the block's row range), so the sums stay in registers; one atomic per column and block.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) ew_colsum_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int rows,
                                                        int cols, float* __restrict__ csum, int rows_per_block, uint32_t thr, float inv_keep,
                                                        uint64_t seed, const unsigned long long* __restrict__ salt) {
  ts::pdl_enter();
  __shared__ float red[8][256 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  if (thr) seed = salted_seed(seed, salt);
  const DropKey key = flat_drop_key(seed, thr);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < cols) {
    for (int r = r0 + w; r < r1; r += 16) {
      float va[2][8], vb[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rr = r + 8 * u;
        if (rr < r1) {
          load8<T>(a + (long long)rr * cols + col, va[u]);
          if (MODE == 1) load8<T>(b + (long long)rr * cols + col, vb[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rr = r + 8 * u;
        if (rr >= r1) break;
        const long long e = (long long)rr * cols + col;
        float o[8], ds[8];
        if (thr) dropout_scale8(key, (uint64_t)e, inv_keep, ds);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v = MODE == 1 ? va[u][i] * gelu_grad_t<T>(vb[u][i]) : va[u][i];
          if (thr) v *= ds[i];
          o[i] = v;
        }
        store8<T>(out + e, o);
        round8<T>(o);                       // the sum of what the weight-gradient GEMM will read
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += o[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i][c];
  if (blockIdx.x * 256 + c < cols) atomicAdd(&csum[blockIdx.x * 256 + c], t);
}

// The same pass with its operands streamed through a per-thread cp.async ring (see vec.cuh): (S - 1) stages x 2 rows x (1-2) vectors
// per thread in flight without holding registers; used for the big GELU-backward pass ([B*T, ffn] = 3 x 37 MB per layer).
template <typename T, int MODE>
__global__ void __launch_bounds__(256, 3) ew_colsum_ring_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int rows,
                                                                int cols, float* __restrict__ csum, int rows_per_block, uint32_t thr,
                                                                float inv_keep, uint64_t seed, const unsigned long long* __restrict__ salt) {
  ts::pdl_enter();
  constexpr int NV = sizeof(T) / 2, NSRC = MODE == 1 ? 2 : 1, VPS = 2 * NSRC * NV, S = 4;
  extern __shared__ __align__(16) uint8_t smraw[];
  uint4* stg = reinterpret_cast<uint4*>(smraw);                               // [S][VPS][256]
  float (*red)[256 + 8] = reinterpret_cast<float (*)[256 + 8]>(stg + S * VPS * 256);   // [8][264]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  if (thr) seed = salted_seed(seed, salt);
  const DropKey key = flat_drop_key(seed, thr);
  const bool active = col < cols;
  auto slot_ptr = [&](int slot, int u, int src) { return stg + ((slot * VPS) + (u * NSRC + src) * NV) * 256 + threadIdx.x; };
  auto issue = [&](int r, int slot) {
    if (active) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rr = r + 8 * u;
        const bool ok = rr < r1;
        const long long e = ok ? (long long)rr * cols + col : (long long)col;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          cp_async16_zfill(slot_ptr(slot, u, 0) + v * 256, reinterpret_cast<const uint8_t*>(a + e) + 16 * v, ok ? 16 : 0);
          if (MODE == 1) cp_async16_zfill(slot_ptr(slot, u, 1) + v * 256, reinterpret_cast<const uint8_t*>(b + e) + 16 * v, ok ? 16 : 0);
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int p0 = 0; p0 < S - 1; ++p0) issue(r0 + w + 16 * p0, p0);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  int it = 0;
  for (int r = r0 + w; r < r1; r += 16, ++it) {
    issue(r + 16 * (S - 1), (it + S - 1) % S);
    cp_async_wait<S - 1>();
    if (!active) continue;
    const int slot = it % S;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int rr = r + 8 * u;
      if (rr >= r1) break;
      const long long e = (long long)rr * cols + col;
      float va[8], vb[8], o[8], ds[8];
      Raw8<T> raw;
      load_raw8_smem(slot_ptr(slot, u, 0), 256, raw);
      unpack8(raw, va);
      if (MODE == 1) { load_raw8_smem(slot_ptr(slot, u, 1), 256, raw); unpack8(raw, vb); }
      if (thr) dropout_scale8(key, (uint64_t)e, inv_keep, ds);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = MODE == 1 ? va[i] * gelu_grad_t<T>(vb[i]) : va[i];
        if (thr) v *= ds[i];
        o[i] = v;
      }
      store8<T>(out + e, o);
      round8<T>(o);                       // the sum of what the weight-gradient GEMM will read
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += o[i];
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i][c];
  if (blockIdx.x * 256 + c < cols) atomicAdd(&csum[blockIdx.x * 256 + c], t);
}

template <int MODE>
static int ew_colsum_launch(Ctx* ctx, int dt, const void* a, const void* b, void* out, int rows, int cols, float* csum, float drop,
                            uint64_t seed, cudaStream_t st) {
  TS_REQUIRE(ctx, rows > 0 && cols > 0 && cols % 8 == 0 && csum, TS_ESHAPE, "ew_colsum: rows=%d cols=%d", rows, cols);
  uint32_t thr; float ik;
  drop_params(drop, &thr, &ik);
  const int cb = cdiv(cols, 256);
  int rb = (ctx->num_sms * 4) / cb;
  if (rb < 1) rb = 1;
  int rpb = cdiv(rows, rb);
  rpb = ((rpb + 15) / 16) * 16;
  dim3 grid(cb, cdiv(rows, rpb));
  static const bool direct = getenv("TETHYS_EW_DIRECT") && atoi(getenv("TETHYS_EW_DIRECT")) != 0;
  if (dt == TS_BF16 && MODE == 1 && !direct && (long long)rows * cols >= (1ll << 22)) {   // big passes: cp.async ring
    const size_t smem = (size_t)4 * (2 * 2) * 256 * 16 + sizeof(float) * 8 * (256 + 8);
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(ew_colsum_ring_kernel<bf16, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ts::launch_k(ew_colsum_ring_kernel<bf16, MODE>, grid, 256, smem, st, (const bf16*)a, (const bf16*)b, (bf16*)out, rows, cols, csum, rpb, thr, ik, seed, ctx->d_state);
    TS_LAUNCH_OK(ctx);
    return 0;
  }
  if (dt == TS_F32) ts::launch_k(ew_colsum_kernel<float, MODE>, grid, 256, 0, st, (const float*)a, (const float*)b, (float*)out, rows, cols, csum, rpb, thr, ik, seed, ctx->d_state);
  else if (dt == TS_BF16) ts::launch_k(ew_colsum_kernel<bf16, MODE>, grid, 256, 0, st, (const bf16*)a, (const bf16*)b, (bf16*)out, rows, cols, csum, rpb, thr, ik, seed, ctx->d_state);
  else return set_err(ctx, TS_EDTYPE, "ew_colsum: dtype %d", dt);
  TS_LAUNCH_OK(ctx);
  return 0;
}
int gelu_bwd_colsum(Ctx* ctx, int dt, const void* df, const void* u, void* du, int rows, int cols, float* csum, float drop, uint64_t seed,
                    cudaStream_t st) {
  return ew_colsum_launch<1>(ctx, dt, df, u, du, rows, cols, csum, drop, seed, st);
}
int dropout_colsum(Ctx* ctx, int dt, const void* x, void* y, int rows, int cols, float* csum, float drop, uint64_t seed, cudaStream_t st) {
  return ew_colsum_launch<2>(ctx, dt, x, nullptr, y, rows, cols, csum, drop, seed, st);
}

// [B,R,C] -> [B,C,R] through a 32x32 smem tile (coalesced both ways)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) transpose_kernel(const TI* __restrict__ x, TO* __restrict__ y, int R, int C) {
  ts::pdl_enter();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? to_f<TI>(x[((long long)b * R + r) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < C) y[((long long)b * C + c) * R + r] = from_f<TO>(tile[tx][i]);
  }
}
int transpose_inner(Ctx* ctx, int dt_in, int dt_out, const void* x, void* y, int B, int R, int C, cudaStream_t st) {
  dim3 grid(cdiv(C, 32), cdiv(R, 32), B);
  if (dt_in == TS_F32 && dt_out == TS_F32) ts::launch_k(transpose_kernel<float, float>, grid, 256, 0, st, (const float*)x, (float*)y, R, C);
  else if (dt_in == TS_F32 && dt_out == TS_BF16) ts::launch_k(transpose_kernel<float, bf16>, grid, 256, 0, st, (const float*)x, (bf16*)y, R, C);
  else if (dt_in == TS_BF16 && dt_out == TS_BF16) ts::launch_k(transpose_kernel<bf16, bf16>, grid, 256, 0, st, (const bf16*)x, (bf16*)y, R, C);
  else return set_err(ctx, TS_EDTYPE, "transpose: dtype %d->%d", dt_in, dt_out);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
