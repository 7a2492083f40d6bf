// K10 (unfused form) — row softmax forward/backward over materialised attention scores, with the reference's
// mask semantics (W:147-160, V:348-359). The score/prob buffer is [nbatch*Tq, ld] with ld a multiple of 8.
//   mask_mode 1 = Whisper decoder self-attention: mask = 1 - band_part(ones,-1,0) (W:416-418) fed through
//   scores + (1-mask)*-1e9 (W:152-154): -1e9 lands on j <= i ("anti-causal"), added in fp32 so that the fully
//   masked last row absorbs the scores and becomes uniform.
#include "ops.cuh"
#include "vec.cuh"

namespace ts {

template <typename T, int NCH>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(T* __restrict__ s, long long ld, long long rows, int Tq, int Tk,
                                                          float scale, int mask_mode, uint32_t thr, float inv_keep,
                                                          uint64_t seed, T* __restrict__ pdrop,
                                                          const unsigned long long* __restrict__ salt) {
  ts::pdl_enter();
  if (thr) seed = salted_seed(seed, salt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int qi = (int)(row % Tq);
  T* sr = s + row * ld;
  float v[NCH][8];
  float mx = -INFINITY;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
    if (col < Tk) {
      load8<T>(sr + col, v[ch]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = v[ch][i] * scale;
        if (mask_mode == 1 && col + i <= qi) x = x + (-1e9f);
        if (col + i >= Tk) x = -INFINITY;
        v[ch][i] = x;
        mx = fmaxf(mx, x);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[ch][i] = -INFINITY;
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float e = (v[ch][i] == -INFINITY) ? 0.f : __expf(v[ch][i] - mx);
      v[ch][i] = e;
      sum += e;
    }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
    if (col < Tk) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = v[ch][i] * inv;
      store8<T>(sr + col, o);
      if (pdrop) {
        round8<T>(o);
        float ds[8];
        dropout_scale8(flat_drop_key(seed, thr), (uint64_t)(row * ld + col), inv_keep, ds);   // ld, col multiples of 8
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= ds[i];
        store8<T>(pdrop + row * ld + col, o);
      }
    }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const T* __restrict__ p, T* __restrict__ dp, long long ld,
                                                          long long rows, int Tk, float scale, uint32_t thr, float inv_keep,
                                                          uint64_t seed, const unsigned long long* __restrict__ salt) {
  ts::pdl_enter();
  if (thr) seed = salted_seed(seed, salt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  float pv[NCH][8], dv[NCH][8];
  float dot = 0.f;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
    if (col < Tk) {
      load8<T>(p + row * ld + col, pv[ch]);
      load8<T>(dp + row * ld + col, dv[ch]);
      float ds[8];
      if (thr) dropout_scale8(flat_drop_key(seed, thr), (uint64_t)(row * ld + col), inv_keep, ds);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (col + i >= Tk) { pv[ch][i] = 0.f; dv[ch][i] = 0.f; }
        if (thr) dv[ch][i] *= ds[i];
        dot += pv[ch][i] * dv[ch][i];
      }
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int col = ch * 256 + lane * 8;
    if (col < Tk) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = scale * pv[ch][i] * (dv[ch][i] - dot);
      store8<T>(dp + row * ld + col, o);
    }
  }
}

static inline void drop_params2(float drop, uint32_t* thr, float* inv_keep) {
  if (drop <= 0.f) { *thr = 0; *inv_keep = 1.f; return; }
  double t = (double)drop * 4294967296.0;
  if (t > 4294967295.0) t = 4294967295.0;
  *thr = (uint32_t)t;
  *inv_keep = 1.f / (1.f - drop);
}

template <typename T>
static int softmax_fwd_t(Ctx* ctx, void* s, long long ld, int nbatch, int Tq, int Tk, float scale, int mask_mode,
                         float drop, uint64_t seed, void* p_drop, cudaStream_t st) {
  const long long rows = (long long)nbatch * Tq;
  const int nch = cdiv(Tk, 256);
  uint32_t thr; float ik;
  drop_params2(drop, &thr, &ik);
  T* pd = thr ? (T*)p_drop : nullptr;
  dim3 grid((unsigned)((rows + 7) / 8));
#define SM_CASE(N)                                                                                                   \
  case N: ts::launch_k(softmax_fwd_kernel<T, N>, grid, 256, 0, st, (T*)s, ld, rows, Tq, Tk, scale, mask_mode, thr, ik, seed, pd, ctx->d_state); break;
  switch (nch) {
    SM_CASE(1) SM_CASE(2) SM_CASE(3) SM_CASE(4) SM_CASE(5) SM_CASE(6) SM_CASE(7) SM_CASE(8)
    default: return set_err(ctx, TS_EUNSUPPORTED, "softmax: Tk=%d > 2048 unsupported", Tk);
  }
#undef SM_CASE
  TS_LAUNCH_OK(ctx);
  return 0;
}

int softmax_fwd(Ctx* ctx, int dt, void* s, long long ld, int nbatch, int Tq, int Tk, float scale, int mask_mode, float drop,
                uint64_t seed, void* p_drop, cudaStream_t st) {
  TS_REQUIRE(ctx, ld % 8 == 0 && ld >= Tk, TS_ESHAPE, "softmax: ld=%lld must be a multiple of 8 and >= Tk=%d", ld, Tk);
  TS_REQUIRE(ctx, !(drop > 0.f && !p_drop), TS_EINVAL, "softmax: dropout needs a p_drop buffer");
  if (dt == TS_F32) return softmax_fwd_t<float>(ctx, s, ld, nbatch, Tq, Tk, scale, mask_mode, drop, seed, p_drop, st);
  if (dt == TS_BF16) return softmax_fwd_t<bf16>(ctx, s, ld, nbatch, Tq, Tk, scale, mask_mode, drop, seed, p_drop, st);
  return set_err(ctx, TS_EDTYPE, "softmax: dtype %d", dt);
}

template <typename T>
static int softmax_bwd_t(Ctx* ctx, const void* p, void* dp, long long ld, int nbatch, int Tq, int Tk, float scale,
                         float drop, uint64_t seed, cudaStream_t st) {
  const long long rows = (long long)nbatch * Tq;
  const int nch = cdiv(Tk, 256);
  uint32_t thr; float ik;
  drop_params2(drop, &thr, &ik);
  dim3 grid((unsigned)((rows + 7) / 8));
#define SMB_CASE(N) \
  case N: ts::launch_k(softmax_bwd_kernel<T, N>, grid, 256, 0, st, (const T*)p, (T*)dp, ld, rows, Tk, scale, thr, ik, seed, ctx->d_state); break;
  switch (nch) {
    SMB_CASE(1) SMB_CASE(2) SMB_CASE(3) SMB_CASE(4) SMB_CASE(5) SMB_CASE(6) SMB_CASE(7) SMB_CASE(8)
    default: return set_err(ctx, TS_EUNSUPPORTED, "softmax_bwd: Tk=%d > 2048 unsupported", Tk);
  }
#undef SMB_CASE
  TS_LAUNCH_OK(ctx);
  return 0;
}

int softmax_bwd(Ctx* ctx, int dt, const void* p, void* dp, long long ld, int nbatch, int Tq, int Tk, float scale, float drop,
                uint64_t seed, cudaStream_t st) {
  TS_REQUIRE(ctx, ld % 8 == 0 && ld >= Tk, TS_ESHAPE, "softmax_bwd: ld=%lld", ld);
  if (dt == TS_F32) return softmax_bwd_t<float>(ctx, p, dp, ld, nbatch, Tq, Tk, scale, drop, seed, st);
  if (dt == TS_BF16) return softmax_bwd_t<bf16>(ctx, p, dp, ld, nbatch, Tq, Tk, scale, drop, seed, st);
  return set_err(ctx, TS_EDTYPE, "softmax_bwd: dtype %d", dt);
}

}  // namespace ts
