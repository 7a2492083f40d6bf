// K15 hard vector quantiser (V:604-660) and K17 contrastive loss (V:865-899), warp-level reductions.
//
// VQ bit-exactness: distances are accumulated in fp32 strictly sequentially over the group
// dimension with separate multiply and add roundings (__fmul_rn/__fadd_rn, no FMA contraction), exactly like the
// oracle's `dist = dist + (z-e)*(z-e)` loop; argmin takes the first minimum (tf.argmin, App. A-10).
#include "ops.cuh"
#include "vec.cuh"

namespace ts {

constexpr int VQ_WARPS = 8;

template <typename T>
__global__ void __launch_bounds__(VQ_WARPS * 32) vq_fwd_kernel(const T* __restrict__ z, const float* __restrict__ cb,
                                                               T* __restrict__ q, long long* __restrict__ idx,
                                                               int* __restrict__ hist, int M, int G, int V, int D) {
  ts::pdl_enter();
  extern __shared__ float sm[];
  float* scb = sm;                       // [32][D+1] codebook chunk
  float* sz = sm + 32 * (D + 1);         // [VQ_WARPS][D] frames
  const int g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * VQ_WARPS + warp;
  const bool valid = m < M;
  if (valid)
    for (int j = lane; j < D; j += 32) sz[warp * D + j] = to_f<T>(z[(long long)m * G * D + g * D + j]);
  float best = INFINITY;
  int best_i = 0x7fffffff;
  for (int v0 = 0; v0 < V; v0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * D; i += VQ_WARPS * 32) {
      const int c = i / D, j = i % D;
      scb[c * (D + 1) + j] = (v0 + c < V) ? cb[((long long)g * V + v0 + c) * D + j] : 0.f;
    }
    __syncthreads();
    if (valid && v0 + lane < V) {
      float dist = 0.f;
      const float* e = scb + lane * (D + 1);
      const float* zz = sz + warp * D;
      for (int j = 0; j < D; ++j) {
        const float d = __fsub_rn(zz[j], e[j]);
        dist = __fadd_rn(dist, __fmul_rn(d, d));
      }
      if (dist < best) { best = dist; best_i = v0 + lane; }  // within a lane codes come in increasing order
    }
  }
  // warp argmin, first minimum wins
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  if (!valid) return;
  if (lane == 0) {
    idx[(long long)g * M + m] = best_i;
    atomicAdd(&hist[g * V + best_i], 1);
  }
  for (int j = lane; j < D; j += 32)
    q[(long long)m * G * D + g * D + j] = from_f<T>(cb[((long long)g * V + best_i) * D + j]);
}

// Second layout, for the group widths the presets use (D = 32 / 64 / 128) and a codebook that fits in shared memory: a lane owns one
// FRAME (its D inputs live in registers), the block's 8 warps split the V codes of the group into 8 slices, and a code vector is
// read as a warp-wide broadcast (one LDS.128 per 4 dimensions) — ~0.3 shared loads per subtract/multiply/add triple instead of
// the 2 of the layout above, which made that one LDS-bound. The arithmetic of a distance is unchanged (sequential over j, separate
// sub / mul / add roundings, no FMA) and ties keep the lowest code index (within a slice by strict <, across slices by visiting
// them in index order), so the int64 indices are bit-identical to the kernel above and to the fp32-sequential restatement.
template <typename T, int D>
__global__ void __launch_bounds__(256) vq_fwd_frames_kernel(const T* __restrict__ z, const float* __restrict__ cb, T* __restrict__ q,
                                                            long long* __restrict__ idx, int* __restrict__ hist, int M, int G, int V) {
  ts::pdl_enter();
  extern __shared__ float sm[];
  float* scb = sm;                                   // [V][D]
  float* sbest = sm + (size_t)V * D;                 // [8][32]
  int* sidx = reinterpret_cast<int*>(sbest + 256);   // [8][32], then [32] winners
  const int g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 32, m = m0 + lane;
  const bool valid = m < M;
  for (int i = threadIdx.x * 4; i < V * D; i += 256 * 4)
    *reinterpret_cast<float4*>(scb + i) = *reinterpret_cast<const float4*>(cb + (size_t)g * V * D + i);
  float zr[D];
  if (valid) {
    const T* zp = z + (long long)m * G * D + g * D;
#pragma unroll
    for (int j = 0; j < D; j += 8) {
      float v8[8];
      load8<T>(zp + j, v8);
#pragma unroll
      for (int e = 0; e < 8; ++e) zr[j + e] = v8[e];
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; ++j) zr[j] = 0.f;
  }
  __syncthreads();
  const int per = (V + 7) / 8, c_lo = warp * per, c_hi = min(V, c_lo + per);
  float best = INFINITY;
  int best_i = 0x7fffffff;
  int c = c_lo;
  for (; c + 1 < c_hi; c += 2) {                     // two codes at a time: two independent accumulation chains
    float d0 = 0.f, d1 = 0.f;
    const float4* e0 = reinterpret_cast<const float4*>(scb + (size_t)c * D);
    const float4* e1 = reinterpret_cast<const float4*>(scb + (size_t)(c + 1) * D);
#pragma unroll
    for (int j = 0; j < D / 4; ++j) {
      const float4 a = e0[j], b = e1[j];
      float t;
      t = __fsub_rn(zr[4 * j], a.x); d0 = __fadd_rn(d0, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j], b.x); d1 = __fadd_rn(d1, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j + 1], a.y); d0 = __fadd_rn(d0, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j + 1], b.y); d1 = __fadd_rn(d1, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j + 2], a.z); d0 = __fadd_rn(d0, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j + 2], b.z); d1 = __fadd_rn(d1, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j + 3], a.w); d0 = __fadd_rn(d0, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j + 3], b.w); d1 = __fadd_rn(d1, __fmul_rn(t, t));
    }
    if (d0 < best) { best = d0; best_i = c; }
    if (d1 < best) { best = d1; best_i = c + 1; }
  }
  if (c < c_hi) {
    float d0 = 0.f;
    const float4* e0 = reinterpret_cast<const float4*>(scb + (size_t)c * D);
#pragma unroll
    for (int j = 0; j < D / 4; ++j) {
      const float4 a = e0[j];
      float t;
      t = __fsub_rn(zr[4 * j], a.x); d0 = __fadd_rn(d0, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j + 1], a.y); d0 = __fadd_rn(d0, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j + 2], a.z); d0 = __fadd_rn(d0, __fmul_rn(t, t));
      t = __fsub_rn(zr[4 * j + 3], a.w); d0 = __fadd_rn(d0, __fmul_rn(t, t));
    }
    if (d0 < best) { best = d0; best_i = c; }
  }
  sbest[warp * 32 + lane] = best;
  sidx[warp * 32 + lane] = best_i;
  __syncthreads();
  if (warp == 0) {
    float b = INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < 8; ++w) {                    // slices in increasing code order: the first minimum wins
      const float ob = sbest[w * 32 + lane];
      if (ob < b) { b = ob; bi = sidx[w * 32 + lane]; }
    }
    sidx[256 + lane] = bi;
    if (valid) {
      idx[(long long)g * M + m] = bi;
      atomicAdd(&hist[g * V + bi], 1);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * D; e += 256) {
    const int f = e / D, j = e % D;
    if (m0 + f < M) q[(long long)(m0 + f) * G * D + g * D + j] = from_f<T>(scb[(size_t)sidx[256 + f] * D + j]);
  }
}

template <typename T, int D>
static int vq_fwd_frames(Ctx* ctx, const void* z, const float* codebook, void* q, long long* idx, int* hist, int M, int G, int V,
                         cudaStream_t st) {
  const size_t smem = ((size_t)V * D + 256) * sizeof(float) + (256 + 32) * sizeof(int);
  static bool attr = false;
  if (!attr) { TS_CUDA_OK(ctx, cudaFuncSetAttribute(vq_fwd_frames_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
  dim3 grid(cdiv(M, 32), G);
  ts::launch_k(vq_fwd_frames_kernel<T, D>, grid, 256, smem, st, (const T*)z, codebook, (T*)q, idx, hist, M, G, V);
  TS_LAUNCH_OK(ctx);
  return 0;
}

int vq_fwd(Ctx* ctx, int dt, const void* z, const float* codebook, void* q, long long* idx, int* hist, int M, int G, int V,
           int D, cudaStream_t st) {
  // frames-in-registers layout when the group's codebook fits in shared memory and rows are 16-byte aligned
  if (((size_t)V * D + 256) * 4 + 288 * 4 <= 200 * 1024 && (V * D) % 4 == 0 && ((size_t)G * D * (dt == TS_F32 ? 4 : 2)) % 16 == 0 &&
      (reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(codebook) & 15) == 0) {
    if (dt == TS_BF16) {
      if (D == 128) return vq_fwd_frames<bf16, 128>(ctx, z, codebook, q, idx, hist, M, G, V, st);
      if (D == 64) return vq_fwd_frames<bf16, 64>(ctx, z, codebook, q, idx, hist, M, G, V, st);
      if (D == 32) return vq_fwd_frames<bf16, 32>(ctx, z, codebook, q, idx, hist, M, G, V, st);
    } else if (dt == TS_F32) {
      if (D == 128) return vq_fwd_frames<float, 128>(ctx, z, codebook, q, idx, hist, M, G, V, st);
      if (D == 64) return vq_fwd_frames<float, 64>(ctx, z, codebook, q, idx, hist, M, G, V, st);
      if (D == 32) return vq_fwd_frames<float, 32>(ctx, z, codebook, q, idx, hist, M, G, V, st);
    }
  }
  const size_t smem = (32 * (D + 1) + VQ_WARPS * D) * sizeof(float);
  TS_REQUIRE(ctx, smem <= 200 * 1024, TS_EUNSUPPORTED, "vq: group dim %d too large", D);
  dim3 grid(cdiv(M, VQ_WARPS), G);
  if (dt == TS_F32) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(vq_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ts::launch_k(vq_fwd_kernel<float>, grid, VQ_WARPS * 32, smem, st, (const float*)z, codebook, (float*)q, idx, hist, M, G, V, D);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(vq_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ts::launch_k(vq_fwd_kernel<bf16>, grid, VQ_WARPS * 32, smem, st, (const bf16*)z, codebook, (bf16*)q, idx, hist, M, G, V, D);
  }
  TS_LAUNCH_OK(ctx);
  return 0;
}

// perplexity = mean_g exp(-sum_v p*log(p+1e-10)), p = clip(count/M, 1e-10, 1)   (V:653-660)
__global__ void vq_perplexity_kernel(const int* __restrict__ hist, float* __restrict__ out, int M, int G, int V) {
  ts::pdl_enter();
  __shared__ float red[32];
  float total = 0.f;
  for (int g = 0; g < G; ++g) {
    float s = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      float p = (float)hist[g * V + v] / (float)M;
      p = fminf(fmaxf(p, 1e-10f), 1.0f);
      s += p * logf(p + 1e-10f);
    }
    s = block_sum(s, red);
    total += expf(-s);
  }
  if (threadIdx.x == 0) out[0] = total / G;
}
int vq_perplexity(Ctx* ctx, const int* hist, float* perplexity, int M, int G, int V, cudaStream_t st) {
  ts::launch_k(vq_perplexity_kernel, 1, 256, 0, st, hist, perplexity, M, G, V);
  TS_LAUNCH_OK(ctx);
  return 0;
}

template <typename T>
__global__ void vq_bwd_kernel(const T* __restrict__ dq, const long long* __restrict__ idx, float* __restrict__ dcb, int M,
                              int G, int V, int D) {
  ts::pdl_enter();
  const long long total = (long long)M * G * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % D);
    const long long r = i / D;
    const int g = (int)(r % G);
    const long long m = r / G;
    const long long code = idx[(long long)g * M + m];
    atomicAdd(&dcb[((long long)g * V + code) * D + j], to_f<T>(dq[i]));
  }
}
int vq_bwd(Ctx* ctx, int dt, const void* dq, const long long* idx, float* dcodebook, int M, int G, int V, int D,
           cudaStream_t st) {
  const long long total = (long long)M * G * D;
  const int grid = (int)min((total + 255) / 256, (long long)ctx->num_sms * 16);
  if (dt == TS_F32) ts::launch_k(vq_bwd_kernel<float>, grid, 256, 0, st, (const float*)dq, idx, dcodebook, M, G, V, D);
  else ts::launch_k(vq_bwd_kernel<bf16>, grid, 256, 0, st, (const bf16*)dq, idx, dcodebook, M, G, V, D);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// ---- contrastive loss over the all-pairs similarity matrix ------------------------------------------------
constexpr int CL_WARPS = 4;

template <typename T>
__global__ void __launch_bounds__(CL_WARPS * 32) contrastive_kernel(const float* __restrict__ S, long long ld,
                                                                    const int* __restrict__ neg, long long neg_bs,
                                                                    long long neg_ts, T* __restrict__ dS, long long ld_ds,
                                                                    float* __restrict__ logits, float* __restrict__ loss_sum,
                                                                    int B, int T_, int K, float inv_temp, float gscale) {
  ts::pdl_enter();
  extern __shared__ float sm[];  // [CL_WARPS][ld_ds] gradient rows
  __shared__ float sloss[CL_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * CL_WARPS + warp;  // row = b*T + t
  float* grow = sm + (long long)warp * ld_ds;
  float myloss = 0.f;
  if (r < (long long)B * T_) {
    const int b = (int)(r / T_), t = (int)(r % T_);
    const float* srow = S + r * ld;
    const int* nrow = neg + b * neg_bs + t * neg_ts;
    for (int j = lane; j < ld_ds; j += 32) grow[j] = 0.f;
    // logits: index 0 = positive (column t), 1..K = negatives
    float mx = -INFINITY;
    for (int k = lane; k <= K; k += 32) {
      const int col = (k == 0) ? t : nrow[k - 1];
      mx = fmaxf(mx, srow[col] * inv_temp);
    }
    mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k <= K; k += 32) {
      const int col = (k == 0) ? t : nrow[k - 1];
      se += __expf(srow[col] * inv_temp - mx);
    }
    se = warp_sum(se);
    const float lse = mx + logf(se);
    __syncwarp();
    for (int k = lane; k <= K; k += 32) {
      const int col = (k == 0) ? t : nrow[k - 1];
      const float l = srow[col] * inv_temp;
      if (logits) logits[r * (K + 1) + k] = l;
      const float pk = __expf(l - lse);
      const float gl = (pk - (k == 0 ? 1.f : 0.f)) * gscale * inv_temp;
      atomicAdd(&grow[col], gl);
    }
    if (lane == 0) myloss = lse - srow[t] * inv_temp;
    __syncwarp();
    for (int j = lane; j < ld_ds; j += 32) dS[r * ld_ds + j] = from_f<T>(grow[j]);
  }
  if (lane == 0) sloss[warp] = myloss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < CL_WARPS; ++i) s += sloss[i];
    atomicAdd(loss_sum, s);
  }
}

int contrastive_fwd_bwd(Ctx* ctx, int dt, const float* S, long long ld, const int* neg, long long neg_bs, long long neg_ts,
                        void* dS, long long ld_ds, float* logits, float* loss_sum, int B, int T_, int K, float temp,
                        float grad_scale, cudaStream_t st) {
  const size_t smem = (size_t)CL_WARPS * ld_ds * sizeof(float);
  TS_REQUIRE(ctx, smem <= 160 * 1024, TS_EUNSUPPORTED, "contrastive: T=%d too long", T_);
  const long long rows = (long long)B * T_;
  dim3 grid((unsigned)((rows + CL_WARPS - 1) / CL_WARPS));
  if (dt == TS_F32) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(contrastive_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ts::launch_k(contrastive_kernel<float>, grid, CL_WARPS * 32, smem, st, S, ld, neg, neg_bs, neg_ts, (float*)dS, ld_ds, logits,
                                                                 loss_sum, B, T_, K, 1.f / temp, grad_scale);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(contrastive_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ts::launch_k(contrastive_kernel<bf16>, grid, CL_WARPS * 32, smem, st, S, ld, neg, neg_bs, neg_ts, (bf16*)dS, ld_ds, logits, loss_sum,
                                                                B, T_, K, 1.f / temp, grad_scale);
  }
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
