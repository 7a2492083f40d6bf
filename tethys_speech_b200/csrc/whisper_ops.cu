// Whisper-specific bandwidth-bound kernels: K13 token embedding gather / scatter-add (W:382, W:405-411), sinusoid
// positional add (W:66-69, W:339, W:408), K14 shifted sparse softmax cross-entropy over the 51 865-class vocabulary
// (W:585-600, double label shift of App. C-2) and the GELU backward of the conv stem with the strided-conv col2im
// fused on load (W:332-336).
#include "ops.cuh"
#include "vec.cuh"

namespace ts {

// ---- cross entropy ------------------------------------------------------------------------------------------
// one block per (b, s) row. Rows s < S-1 are scored against labels[b, s+1]; row S-1 has no target (logits[:, :-1]).
template <typename T>
__global__ void __launch_bounds__(512) ce_kernel(const T* logits, T* dlogits /* may alias logits */, long long ldv,
                                                 const int* __restrict__ labels, float* __restrict__ loss_sum, int S, int V,
                                                 float gscale) {
  ts::pdl_enter();
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const int b = (int)(row / S), s = (int)(row % S);
  const T* lr = logits + row * ldv;
  T* dr = dlogits + row * ldv;
  if (s == S - 1) {
    for (long long j = threadIdx.x; j < ldv; j += blockDim.x) dr[j] = from_f<T>(0.f);
    return;
  }
  const int target = labels[b * S + s + 1];
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < V; j += blockDim.x) mx = fmaxf(mx, to_f<T>(lr[j]));
  mx = block_max(mx, red);
  float se = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) se += __expf(to_f<T>(lr[j]) - mx);
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  if (threadIdx.x == 0) atomicAdd(loss_sum, lse - to_f<T>(lr[target]));
  __syncthreads();  // lr[target] read before the in-place overwrite below
  for (long long j = threadIdx.x; j < ldv; j += blockDim.x) {
    float g = 0.f;
    if (j < V) g = (__expf(to_f<T>(lr[j]) - lse) - (j == target ? 1.f : 0.f)) * gscale;
    dr[j] = from_f<T>(g);
  }
}

// The same with the row held in shared memory: ONE pass over HBM in 16-byte vectors (the row — 51 872 bf16 = 104 KB for Whisper's
// vocabulary — is staged while its maximum is taken), the sum of exponentials and the gradient are computed from the staged copy and
// the gradient row is written in 16-byte vectors. (The scalar kernel above made three passes of 2-byte loads: 107 us for 82 MB.)
template <typename T>
__global__ void __launch_bounds__(512) ce_row_smem_kernel(const T* logits, T* dlogits /* may alias logits */, long long ldv,
                                                          const int* __restrict__ labels, float* __restrict__ loss_sum, int S, int V,
                                                          float gscale) {
  ts::pdl_enter();
  extern __shared__ __align__(16) uint8_t ce_smem[];
  __shared__ float red[32];
  T* srow = reinterpret_cast<T*>(ce_smem);
  const long long row = blockIdx.x;
  const int b = (int)(row / S), s = (int)(row % S);
  const T* lr = logits + row * ldv;
  T* dr = dlogits + row * ldv;
  const int nvec = (int)(ldv / 8);
  if (s == S - 1) {
    float z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = 0.f;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) store8<T>(dr + (long long)v * 8, z);
    return;
  }
  const int target = labels[b * S + s + 1];
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float x[8];
    load8<T>(lr + (long long)v * 8, x);
    store8<T>(srow + v * 8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (v * 8 + i < V) mx = fmaxf(mx, x[i]);
  }
  mx = block_max(mx, red);     // (block_* synchronise: the staged row is complete for every thread after this)
  float se = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float x[8];
    load8<T>(srow + v * 8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (v * 8 + i < V) se += __expf(x[i] - mx);
  }
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  if (threadIdx.x == 0) atomicAdd(loss_sum, lse - to_f<T>(srow[target]));
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float x[8], g[8];
    load8<T>(srow + v * 8, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = v * 8 + i;
      g[i] = j < V ? (__expf(x[i] - lse) - (j == target ? 1.f : 0.f)) * gscale : 0.f;
    }
    store8<T>(dr + (long long)v * 8, g);
  }
}

template <typename T>
static int ce_launch(Ctx* ctx, const void* logits, void* dlogits, long long ldv, const int* labels, float* loss_sum, int B, int S, int V,
                     float gs, cudaStream_t st) {
  const size_t smem = (size_t)ldv * sizeof(T);
  const bool al = ldv % 8 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0;
  if (al && smem <= 200 * 1024) {
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(ce_row_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ts::launch_k(ce_row_smem_kernel<T>, B * S, 512, smem, st, (const T*)logits, (T*)dlogits, ldv, labels, loss_sum, S, V, gs);
  } else {
    ts::launch_k(ce_kernel<T>, B * S, 512, 0, st, (const T*)logits, (T*)dlogits, ldv, labels, loss_sum, S, V, gs);
  }
  TS_LAUNCH_OK(ctx);
  return 0;
}

int ce_fwd_bwd(Ctx* ctx, int dt, const void* logits, void* dlogits, long long ldv, const int* labels, float* loss_sum, int B,
               int S, int V, float grad_scale, cudaStream_t st) {
  const float gs = grad_scale / (float)(B * (S - 1));
  if (dt == TS_F32) return ce_launch<float>(ctx, logits, dlogits, ldv, labels, loss_sum, B, S, V, gs, st);
  if (dt == TS_BF16) return ce_launch<bf16>(ctx, logits, dlogits, ldv, labels, loss_sum, B, S, V, gs, st);
  return set_err(ctx, TS_EDTYPE, "ce: dtype %d", dt);
}

// ---- embedding -------------------------------------------------------------------------------------------------
// decoder_input_ids = pad(labels[:, :-1], left, start_token)   (W:559-563)
__device__ __forceinline__ int dec_id(const int* labels, int b, int s, long long ld, int start_token) {
  return s == 0 ? start_token : labels[b * ld + s - 1];
}

template <typename T>
__global__ void embed_fwd_kernel(const T* __restrict__ table, const int* __restrict__ labels, const float* __restrict__ pe,
                                 T* __restrict__ out, int S, long long label_ld, int D, int start_token, uint32_t thr, float inv_keep,
                                 uint64_t seed, const unsigned long long* __restrict__ salt) {
  ts::pdl_enter();
  if (thr) seed = salted_seed(seed, salt);
  const long long row = blockIdx.x;
  const int b = (int)(row / S), s = (int)(row % S);
  const long long id = dec_id(labels, b, s, label_ld, start_token);
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float v = to_f<T>(table[id * D + j]) + pe[(long long)s * D + j];
    if (thr) v *= dropout_scale(seed, (uint64_t)(row * D + j), thr, inv_keep);
    out[row * D + j] = from_f<T>(v);
  }
}
template <typename T>
__global__ void embed_bwd_kernel(const T* __restrict__ dout, const int* __restrict__ labels, float* __restrict__ dtable, int S,
                                 int D, int start_token, uint32_t thr, float inv_keep, uint64_t seed,
                                 const unsigned long long* __restrict__ salt) {
  ts::pdl_enter();
  if (thr) seed = salted_seed(seed, salt);
  const long long row = blockIdx.x;
  const int b = (int)(row / S), s = (int)(row % S);
  const long long id = dec_id(labels, b, s, S, start_token);
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float g = to_f<T>(dout[row * D + j]);
    if (thr) g *= dropout_scale(seed, (uint64_t)(row * D + j), thr, inv_keep);
    atomicAdd(&dtable[id * D + j], g);
  }
}

static inline void drop3(float drop, uint32_t* thr, float* ik) {
  if (drop <= 0.f) { *thr = 0; *ik = 1.f; return; }
  double t = (double)drop * 4294967296.0;
  *thr = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
  *ik = 1.f / (1.f - drop);
}

int embed_fwd(Ctx* ctx, int dt, const void* table, const int* labels, long long label_ld, const float* pe, void* out, int B, int S, int D,
              int start_token, float drop, uint64_t seed, cudaStream_t st) {
  uint32_t thr; float ik;
  drop3(drop, &thr, &ik);
  if (dt == TS_F32) ts::launch_k(embed_fwd_kernel<float>, B * S, 256, 0, st, (const float*)table, labels, pe, (float*)out, S, label_ld, D, start_token, thr, ik, seed, ctx->d_state);
  else ts::launch_k(embed_fwd_kernel<bf16>, B * S, 256, 0, st, (const bf16*)table, labels, pe, (bf16*)out, S, label_ld, D, start_token, thr, ik, seed, ctx->d_state);
  TS_LAUNCH_OK(ctx);
  return 0;
}
int embed_bwd(Ctx* ctx, int dt, const void* dout, const int* labels, float* dtable, int B, int S, int D, int start_token,
              float drop, uint64_t seed, cudaStream_t st) {
  uint32_t thr; float ik;
  drop3(drop, &thr, &ik);
  if (dt == TS_F32) ts::launch_k(embed_bwd_kernel<float>, B * S, 256, 0, st, (const float*)dout, labels, dtable, S, D, start_token, thr, ik, seed, ctx->d_state);
  else ts::launch_k(embed_bwd_kernel<bf16>, B * S, 256, 0, st, (const bf16*)dout, labels, dtable, S, D, start_token, thr, ik, seed, ctx->d_state);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// ---- y[b,t,:] = dropout(x[b,t,:] + pe[t,:]) with x stored rpb_in rows per batch, y dense [B,T,D] ------------------
template <typename T>
__global__ void add_pe_kernel(const T* __restrict__ x, long long rpb_in, const float* __restrict__ pe, T* __restrict__ y, int T_,
                              int D, uint32_t thr, float inv_keep, uint64_t seed, long long total8,
                              const unsigned long long* __restrict__ salt) {
  ts::pdl_enter();
  if (thr) seed = salted_seed(seed, salt);
  const int d8 = D / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % d8) * 8;
    const long long r = i / d8;
    const int t = (int)(r % T_);
    const long long b = r / T_;
    float v[8], p[8];
    load8<T>(x + (b * rpb_in + t) * D + c, v);
    load8<float>(pe + (long long)t * D + c, p);
    float ds[8];
    if (thr) dropout_scale8(flat_drop_key(seed, thr), (uint64_t)(r * D + c), inv_keep, ds);   // D % 8 == 0: 8-aligned
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] += p[j];
      if (thr) v[j] *= ds[j];
    }
    store8<T>(y + r * D + c, v);
  }
}
int add_pe_rows(Ctx* ctx, int dt, const void* x, long long rpb_in, const float* pe, void* y, int B, int T_, int D, float drop,
                uint64_t seed, cudaStream_t st) {
  TS_REQUIRE(ctx, D % 8 == 0, TS_ESHAPE, "add_pe: D=%d", D);
  uint32_t thr; float ik;
  drop3(drop, &thr, &ik);
  const long long total8 = (long long)B * T_ * (D / 8);
  const int grid = (int)min((total8 + 255) / 256, (long long)ctx->num_sms * 16);
  if (dt == TS_F32) ts::launch_k(add_pe_kernel<float>, grid, 256, 0, st, (const float*)x, rpb_in, pe, (float*)y, T_, D, thr, ik, seed, total8, ctx->d_state);
  else ts::launch_k(add_pe_kernel<bf16>, grid, 256, 0, st, (const bf16*)x, rpb_in, pe, (bf16*)y, T_, D, thr, ik, seed, total8, ctx->d_state);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// ---- du[b,t,:] = da[b,t,:] * gelu'(u[b,t,:]); da dense [B,T,C] or col2im of the next strided conv's dcol -----------
template <typename T>
__global__ void gelu_bwd_rows_kernel(const T* __restrict__ da, long long da_rpb, Col2imSrc col, const T* __restrict__ u,
                                     T* __restrict__ du, long long rpb, int T_, int C, long long total8) {
  ts::pdl_enter();
  const int c8 = C / 8;
  const T* dcol = reinterpret_cast<const T*>(col.dcol);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    const long long r = i / c8;
    const int t = (int)(r % rpb);
    const long long b = r / rpb;
    float o[8];
    if (t < T_) {
      float d[8], uv[8];
      if (dcol) {
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = 0.f;
        for (int j = 0; j < col.k; ++j) {
          const int q = t + col.left - j;
          if (q >= 0 && (q % col.s) == 0 && (q / col.s) < col.t_next) {
            float v[8];
            load8<T>(dcol + (b * col.rows_per_batch + q / col.s) * ((long long)col.k * C) + (long long)j * C + c, v);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) d[jj] += v[jj];
          }
        }
      } else {
        load8<T>(da + (b * da_rpb + t) * C + c, d);
      }
      load8<T>(u + r * C + c, uv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = d[j] * gelu_grad_t<T>(uv[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
    store8<T>(du + r * C + c, o);
  }
}
int gelu_bwd_rows(Ctx* ctx, int dt, const void* da, long long da_rpb, const Col2imSrc* col, const void* u, void* du, long long rpb,
                  int B, int T_, int C, cudaStream_t st) {
  TS_REQUIRE(ctx, C % 8 == 0, TS_ESHAPE, "gelu_bwd_rows: C=%d", C);
  Col2imSrc c;
  if (col) c = *col; else { c.dcol = nullptr; c.rows_per_batch = 0; c.t_next = 0; c.k = 0; c.s = 1; c.left = 0; }
  const long long total8 = (long long)B * rpb * (C / 8);
  const int grid = (int)min((total8 + 255) / 256, (long long)ctx->num_sms * 16);
  if (dt == TS_F32) ts::launch_k(gelu_bwd_rows_kernel<float>, grid, 256, 0, st, (const float*)da, da_rpb, c, (const float*)u, (float*)du, rpb, T_, C, total8);
  else ts::launch_k(gelu_bwd_rows_kernel<bf16>, grid, 256, 0, st, (const bf16*)da, da_rpb, c, (const bf16*)u, (bf16*)du, rpb, T_, C, total8);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// ---- zero rows [row_from, rpb) of every batch block ---------------------------------------------------------------
template <typename T>
__global__ void zero_rows_kernel(T* __restrict__ x, long long rpb, int row_from, int C, long long total) {
  ts::pdl_enter();
  const int nz = (int)(rpb - row_from);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int t = row_from + (int)(r % nz);
    const long long b = r / nz;
    x[(b * rpb + t) * C + c] = from_f<T>(0.f);
  }
}
int zero_rows(Ctx* ctx, int dt, void* x, long long rpb, int row_from, int B, int C, cudaStream_t st) {
  if (row_from >= rpb) return 0;
  const long long total = (long long)B * (rpb - row_from) * C;
  const int grid = (int)min((total + 255) / 256, (long long)ctx->num_sms * 8);
  if (dt == TS_F32) ts::launch_k(zero_rows_kernel<float>, grid, 256, 0, st, (float*)x, rpb, row_from, C, total);
  else ts::launch_k(zero_rows_kernel<bf16>, grid, 256, 0, st, (bf16*)x, rpb, row_from, C, total);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
