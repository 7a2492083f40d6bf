// Span masking utilities: apply_time_mask / apply_feature_mask (V:1073-1095, V:1098-1120). The reference defines them but
// never calls them; they are provided as standalone operators. Integer part: every span start
// is dilated to the right by mask_length positions (expanded[t] = OR_{i < mask_length} start[t - i]); float part:
// y = x * (1 - expanded). The starts are an input (uint8), drawn by the caller (tf.random.uniform(...) < mask_prob there).
#include "ops.cuh"
#include "vec.cuh"

namespace ts {

// expanded [B, L] (float 0/1) from starts [B, L] (uint8)
__global__ void span_expand_kernel(const unsigned char* __restrict__ start, float* __restrict__ expanded, int L, int mask_length,
                                   long long total) {
  ts::pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int t = (int)(i % L);
  const unsigned char* row = start + (i - t);
  int hit = 0;
  const int lo = max(0, t - mask_length + 1);
  for (int s = lo; s <= t; ++s) hit |= row[s];
  expanded[i] = hit ? 1.f : 0.f;
}

// y[b, t, h] = x[b, t, h] * (1 - m), m = expanded[b, t] (axis 1) or expanded[b, h] (axis 2); 8 elements per thread
template <typename T>
__global__ void __launch_bounds__(256) span_apply_kernel(const T* __restrict__ x, const float* __restrict__ expanded, T* __restrict__ y,
                                                         int T_, int H, int axis, long long nvec) {
  ts::pdl_enter();
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvec) return;
  const long long e = v * 8;
  const int h = (int)(e % H);
  const long long bt = e / H;
  float a[8];
  load8<T>(x + e, a);
  if (axis == 1) {
    const float keep = 1.f - expanded[bt];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] *= keep;
  } else {
    const float* m = expanded + (bt / T_) * H + h;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] *= 1.f - m[i];
  }
  store8<T>(y + e, a);
}

int span_mask_apply(Ctx* ctx, int dt, const void* x, const unsigned char* start, void* y, float* expanded, int B, int T_, int H,
                    int axis, int mask_length, cudaStream_t st) {
  TS_REQUIRE(ctx, axis == 1 || axis == 2, TS_EINVAL, "span_mask: axis must be 1 (time) or 2 (feature)");
  TS_REQUIRE(ctx, B > 0 && T_ > 0 && H > 0 && H % 8 == 0 && mask_length >= 1, TS_ESHAPE, "span_mask: bad shape (H must be a multiple of 8)");
  TS_REQUIRE(ctx, x && start && y && expanded, TS_EINVAL, "span_mask: null pointer");
  const int L = axis == 1 ? T_ : H;
  const long long total = (long long)B * L;
  ts::launch_k(span_expand_kernel, cdiv(total, 256), 256, 0, st, start, expanded, L, mask_length, total);
  TS_LAUNCH_OK(ctx);
  const long long nvec = (long long)B * T_ * H / 8;
  if (dt == TS_F32) ts::launch_k(span_apply_kernel<float>, cdiv(nvec, 256), 256, 0, st, (const float*)x, expanded, (float*)y, T_, H, axis, nvec);
  else if (dt == TS_BF16) ts::launch_k(span_apply_kernel<bf16>, cdiv(nvec, 256), 256, 0, st, (const bf16*)x, expanded, (bf16*)y, T_, H, axis, nvec);
  else return set_err(ctx, TS_EDTYPE, "span_mask: dtype %d", dt);
  TS_LAUNCH_OK(ctx);
  return 0;
}

// ---- negative sample positions of the contrastive loss (V:907-937) — integer work, bit-exact -----------------------------------
// Per batch row: tf.nn.top_k(-float(r), k) over T uniform ints r = the positions of the k smallest values, smallest first, ties
// by lower index (top_k is stable on equal keys). One block per row: every element's rank = #{smaller} + #{equal, earlier index}
// counted against the row held in smem (T^2 compares, T <= a few thousand); rank < k writes out[rank]; the list is tiled to
// `num_neg` entries when k = min(num_neg, T - 1) < num_neg (V:925-931).
__global__ void __launch_bounds__(256) negatives_kernel(const int* __restrict__ r, int T_, int k, int num_neg, int* __restrict__ out) {
  ts::pdl_enter();
  extern __shared__ int srow[];
  const int* row = r + (long long)blockIdx.x * T_;
  int* o = out + (long long)blockIdx.x * num_neg;
  for (int i = threadIdx.x; i < T_; i += blockDim.x) srow[i] = row[i];
  __syncthreads();
  for (int i = threadIdx.x; i < T_; i += blockDim.x) {
    const int v = srow[i];
    int rank = 0;
    for (int j = 0; j < T_; ++j) rank += (srow[j] < v) || (srow[j] == v && j < i);
    if (rank < k)
      for (int q = rank; q < num_neg; q += k) o[q] = i;
  }
}

int sample_negatives(Ctx* ctx, const int* rand, int B, int T_, int num_neg, int* out, cudaStream_t st) {
  TS_REQUIRE(ctx, rand && out && B > 0 && T_ > 0 && num_neg > 0, TS_EINVAL, "sample_negatives: bad arguments");
  TS_REQUIRE(ctx, T_ <= 48 * 1024, TS_ESHAPE, "sample_negatives: T = %d rows do not fit in shared memory", T_);
  int k = num_neg < T_ - 1 ? num_neg : T_ - 1;
  if (k < 1) k = 1;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(negatives_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 * 4); attr = true; }
  ts::launch_k(negatives_kernel, B, 256, (size_t)T_ * 4, st, rand, T_, k, num_neg, out);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts

extern "C" int ts_w2v_sample_negatives(ts_ctx* ctx, const int32_t* random_ints, int batch, int t, int num_negatives, int32_t* out,
                                       void* stream) {
  if (!ctx) return TS_EINVAL;
  return ts::sample_negatives(reinterpret_cast<ts::Ctx*>(ctx), random_ints, batch, t, num_negatives, out, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ts_span_mask_apply(ts_ctx* ctx, int dtype, const void* x, const unsigned char* start_mask, void* y, float* expanded_mask,
                                  int batch, int t, int h, int axis, int mask_length, void* stream) {
  if (!ctx) return TS_EINVAL;
  return ts::span_mask_apply(reinterpret_cast<ts::Ctx*>(ctx), dtype, x, start_mask, y, expanded_mask, batch, t, h, axis, mask_length,
                             reinterpret_cast<cudaStream_t>(stream));
}
