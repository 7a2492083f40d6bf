// K19 gradient-norm clipping (tf.clip_by_global_norm V:1243, optimizer clipnorm V:1274) and K20 the Keras-2.10
// legacy Adam update (W:901, V:1271-1275), multi-tensor over the flat parameter arena.
// One launch covers every variable: the arena is cut into work items (<= 64K elements of one segment each).
// HBM-bound: 28 B/param fp32 (+2 B/param for the bf16 compute copy written in the same pass).
#include <vector>
#include "ops.cuh"

namespace ts {

// Work items are built ONCE per optimizer object (ts_optim_create) and owned by it: no process-global cache, nothing to evict
// under a captured CUDA graph, no host synchronisation on the step path.
std::vector<WorkItem> build_work_items(const std::vector<Segment>& segs) {
  std::vector<WorkItem> items;
  const long long chunk = 65536;
  for (int s = 0; s < (int)segs.size(); ++s) {
    const Segment& sg = segs[s];
    if (sg.rows == 1) {
      for (long long c0 = 0; c0 < sg.cols; c0 += chunk)
        items.push_back({sg.offset + c0, sg.ld, 1, (int)std::min<long long>(chunk, sg.cols - c0), s, 0});
    } else {
      const int rpi = (int)std::max<long long>(1, chunk / sg.cols);
      for (int r0 = 0; r0 < sg.rows; r0 += rpi)
        items.push_back({sg.offset + (long long)r0 * sg.ld, sg.ld, std::min(rpi, sg.rows - r0), sg.cols, s, 0});
    }
  }
  return items;
}

__global__ void zero_f32_kernel(float* p, int n) {
  ts::pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

// gradient arena element type GT: float (the arena the backward kernels write) or bf16 (the all-reduced gradient bucket, read as is)
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4(const bf16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float ldg1(const float* p) { return *p; }
__device__ __forceinline__ float ldg1(const bf16* p) { return __bfloat162float(*p); }

template <typename GT>
__global__ void __launch_bounds__(256) sumsq_kernel(const GT* __restrict__ g, const WorkItem* __restrict__ items,
                                                    float* __restrict__ sumsq) {
  ts::pdl_enter();
  __shared__ float red[32];
  const WorkItem it = items[blockIdx.x];
  float s = 0.f;
  const bool vec = ((it.start | it.ld | (long long)it.cols) & 3) == 0;   // arena blocks are 64-element aligned: the usual case
  for (int r = 0; r < it.rows; ++r) {
    const GT* p = g + it.start + (long long)r * it.ld;
    if (vec) {
      for (int c = threadIdx.x * 4; c < it.cols; c += 1024) {
        const float4 v = ldg4(p + c);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    } else {
      for (int c = threadIdx.x; c < it.cols; c += 256) { const float v = ldg1(p + c); s += v * v; }
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(&sumsq[it.seg], s);
}

int grad_sumsq(Ctx* ctx, const void* grads, int grad_dt, const WorkItem* items, int n, int nseg, float* sumsq, cudaStream_t st) {
  ts::launch_k(zero_f32_kernel, cdiv(nseg, 256), 256, 0, st, sumsq, nseg);
  if (grad_dt == TS_BF16) ts::launch_k(sumsq_kernel<bf16>, n, 256, 0, st, (const bf16*)grads, items, sumsq);
  else ts::launch_k(sumsq_kernel<float>, n, 256, 0, st, (const float*)grads, items, sumsq);
  TS_LAUNCH_OK(ctx);
  return 0;
}

__global__ void global_clip_kernel(const float* __restrict__ sumsq, int nseg, float clip, float* scale_out, float* norm_out) {
  ts::pdl_enter();
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < nseg; i += blockDim.x) s += sumsq[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float n = sqrtf(s);
    if (norm_out) norm_out[0] = n;
    scale_out[0] = clip / fmaxf(n, clip);   // tf.clip_by_global_norm: g * clip / max(norm, clip)
  }
}
int global_clip_scale(Ctx* ctx, const float* sumsq, int nseg, float clip, float* scale_out, float* norm_out, cudaStream_t st) {
  ts::launch_k(global_clip_kernel, 1, 256, 0, st, sumsq, nseg, clip, scale_out, norm_out);
  TS_LAUNCH_OK(ctx);
  return 0;
}

template <typename GT>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const GT* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, bf16* __restrict__ p16,
                                                   const WorkItem* __restrict__ items, float lr_t, float omb1, float omb2,
                                                   float eps, float clipnorm, const float* __restrict__ pre_scale,
                                                   const float* __restrict__ sumsq, float lr,
                                                   const unsigned long long* __restrict__ step_state) {
  ts::pdl_enter();
  if (step_state) {  // CUDA-graph mode: the step count lives on the device (Ctx::d_state[1]); same formula as the host path
    const double t = (double)step_state[1];
    lr_t = (float)((double)lr * sqrt(1.0 - pow(1.0 - (double)omb2, t)) / (1.0 - pow(1.0 - (double)omb1, t)));
  }
  const WorkItem it = items[blockIdx.x];
  float sc = pre_scale ? pre_scale[0] : 1.f;
  if (clipnorm > 0.f && sumsq) {
    const float n = sqrtf(sumsq[it.seg]) * sc;
    sc *= clipnorm / fmaxf(n, clipnorm);
  }
  const bool vec = ((it.start | it.ld | (long long)it.cols) & 3) == 0;
  for (int r = 0; r < it.rows; ++r) {
    const long long base = it.start + (long long)r * it.ld;
    if (vec) {  // 16-byte accesses: 4 parameters per thread and iteration
      for (int c = threadIdx.x * 4; c < it.cols; c += 1024) {
        const long long i = base + c;
        const float4 g4 = ldg4(g + i);
        float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i);
        float4 p4 = *reinterpret_cast<const float4*>(p + i);
        const float gg[4] = {g4.x * sc, g4.y * sc, g4.z * sc, g4.w * sc};
        float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          mm[k] = mm[k] + (gg[k] - mm[k]) * omb1;
          vv[k] = vv[k] + (gg[k] * gg[k] - vv[k]) * omb2;
          pp[k] = pp[k] - lr_t * mm[k] / (sqrtf(vv[k]) + eps);
        }
        *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
        if (p16) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(pp[0], pp[1]), hi = __floats2bfloat162_rn(pp[2], pp[3]);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&lo);
          u.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(p16 + i) = u;
        }
      }
    } else {
      for (int c = threadIdx.x; c < it.cols; c += 256) {
        const long long i = base + c;
        const float gi = ldg1(g + i) * sc;
        float mi = m[i], vi = v[i];
        mi = mi + (gi - mi) * omb1;
        vi = vi + (gi * gi - vi) * omb2;
        const float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
        m[i] = mi; v[i] = vi; p[i] = pi;
        if (p16) p16[i] = __float2bfloat16_rn(pi);
      }
    }
  }
}

int adam_step(Ctx* ctx, float* params, const void* grads, int grad_dt, float* m, float* v, void* params_bf16, const WorkItem* items, int n,
              const AdamArgs& a, cudaStream_t st) {
  TS_REQUIRE(ctx, a.step >= 0, TS_EINVAL, "adam: step must be >= 1 (or 0 = take it from the device step state)");
  const double lr_t = a.step >= 1 ? (double)a.lr * sqrt(1.0 - pow((double)a.beta2, a.step)) / (1.0 - pow((double)a.beta1, a.step)) : 0.0;
  if (grad_dt == TS_BF16)
    ts::launch_k(adam_kernel<bf16>, n, 256, 0, st, params, (const bf16*)grads, m, v, (bf16*)params_bf16, items, (float)lr_t, 1.f - a.beta1,
                 1.f - a.beta2, a.eps, a.clipnorm, a.pre_scale, a.sumsq, a.lr, a.step >= 1 ? nullptr : ctx->d_state);
  else
    ts::launch_k(adam_kernel<float>, n, 256, 0, st, params, (const float*)grads, m, v, (bf16*)params_bf16, items, (float)lr_t, 1.f - a.beta1,
                 1.f - a.beta2, a.eps, a.clipnorm, a.pre_scale, a.sumsq, a.lr, a.step >= 1 ? nullptr : ctx->d_state);
  TS_LAUNCH_OK(ctx);
  return 0;
}

__global__ void scale_kernel(float* __restrict__ x, long long n, const float* __restrict__ sdev, float shost) {
  ts::pdl_enter();
  const float s = (sdev ? sdev[0] : 1.f) * shost;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= s;
}
int scale_inplace(Ctx* ctx, float* x, long long n, const float* scale_dev, float scale_host, cudaStream_t st) {
  if (n <= 0) return 0;
  const int grid = (int)min((n + 255) / 256, (long long)ctx->num_sms * 16);
  ts::launch_k(scale_kernel, grid, 256, 0, st, x, n, scale_dev, scale_host);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
