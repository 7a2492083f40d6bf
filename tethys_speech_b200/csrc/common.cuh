// Common device/host helpers for the tethys-speech B200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/tethys.h"

namespace ts {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------------
// Context (one per process x GPU). Holds the last error string and device facts.
// ----------------------------------------------------------------------------------------------
struct Ctx {
  int device = 0;
  int num_sms = 148;
  std::string err;
  // driver entry point for cuTensorMapEncodeTiled, resolved lazily (no link-time libcuda dependency,
  // so the library still loads on a CPU-only box for the symbol-export test).
  void* encode_tiled = nullptr;
  int* d_watchdog = nullptr;   // device int: set non-zero by a kernel whose mbarrier wait timed out
  void* tmap_cache = nullptr;  // opaque tensor-map cache (gemm_tc.cu)
  unsigned long long launches = 0;  // kernel-launch sites passed (TS_LAUNCH_OK); reported by ts_launch_count
  unsigned long long simt_downgrades = 0;  // bf16 GEMMs that fell back to the CUDA-core engine (ts_simt_downgrades)
  // device-resident step state {dropout salt, optimizer step}: lets a whole train step be captured in a CUDA graph and still
  // draw fresh dropout masks / use the right Adam bias correction on every replay (ts_step_state_set / _advance)
  unsigned long long* d_state = nullptr;
  void* gemm_trace = nullptr;  // debug: device buffer [grid x 8] of globaltimer stamps written by gemm_tc_kernel (ts_debug_gemm_trace)
};

int set_err(Ctx* c, int code, const char* fmt, ...);

#define TS_CUDA_OK(ctx, expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ts::set_err((ctx), TS_ECUDA, "%s failed: %s (%s:%d)", #expr,                  \
                         cudaGetErrorString(_e), __FILE__, __LINE__);                      \
  } while (0)

#define TS_LAUNCH_OK(ctx)                                                                  \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    (ctx)->launches++;                                                                     \
    if (_e != cudaSuccess)                                                          \
      return ts::set_err((ctx), TS_ECUDA, "kernel launch failed: %s (%s:%d)",              \
                         cudaGetErrorString(_e), __FILE__, __LINE__);                      \
  } while (0)

#define TS_REQUIRE(ctx, cond, code, ...)                                                   \
  do {                                                                                     \
    if (!(cond)) return ts::set_err((ctx), (code), __VA_ARGS__);                           \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL). A train step is ~400 dependent launches of 5-200 us each; between two plain launches
// the GPU drains the grid, flushes, and only then schedules the next grid's CTAs, which then run their own prologue
// (barrier init, TMEM allocation, tensor-map fetch, index math). With the programmatic-stream-serialization attribute the next
// grid's CTAs are scheduled as soon as every CTA of the running grid has passed pdl_enter() and an SM has room; they run
// their prologue and block in griddepcontrol.wait until the previous grid has completed and its memory is visible.
// Contract: EVERY kernel of this library executes pdl_enter() (all threads) before its first global-memory access, so the
// chain is transitive (grid N+1 passes its wait only after grid N completed, which had itself waited for N-1).
// TETHYS_PDL=0 launches without the attribute (griddepcontrol.* are then no-ops).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifndef TS_PDL_MODE
#define TS_PDL_MODE 1
#endif
// mode 0: wait + trigger at entry; 1: wait only (the trigger is implicit at grid completion); 2: wait at entry, the persistent
// tcgen05 kernels trigger from their producer warp once the last operand load has been issued (pdl_tail)
__device__ __forceinline__ void pdl_enter() { pdl_wait(); if (TS_PDL_MODE == 0) pdl_launch_dependents(); }
__device__ __forceinline__ void pdl_tail() { if (TS_PDL_MODE == 2) pdl_launch_dependents(); }

bool pdl_enabled();   // api.cu: reads TETHYS_PDL once

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

// ----------------------------------------------------------------------------------------------
// dtype helpers
// ----------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// exact-erf GELU (reference: tf.keras.activations.gelu default / custom gelu, whisper_dist.py:195,
// wav2vec2_dist.py:132-136) and its derivative.
__device__ __forceinline__ float gelu_f(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// bf16-path variants: erf by Abramowitz-Stegun 7.1.26 (|abs err| < 2e-6 in fp32 arithmetic, far below bf16 resolution)
// with the shared exponential exp(-x^2/2) computed once: ~15 instructions instead of ~35 for erff + expf.
// The fp32 parity mode keeps the exact versions above.
__device__ __forceinline__ void gelu_parts_fast(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.f)));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float w = ax * 0.84932180028801904272f;             // |x| * sqrt(log2(e) / 2): exp(-x^2/2) = 2^(-w^2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-w * w));
  const float erf_abs = fmaf(-p, e, 1.f);                    // erf(|x|/sqrt2)
  cdf = fmaf(copysignf(0.5f, x), erf_abs, 0.5f);
  pdf = 0.39894228040143267794f * e;
}
// Forward-only variant with ONE special-function op: Phi(-|x|) = 0.5 erfc(|x|/sqrt2) = 2^q(|x|), q a degree-6 minimax fit
// of log2(0.5 erfc(t/sqrt2)) on [0, 6] (max |q - log2| = 6.6e-5, i.e. the tail keeps a RELATIVE error of 4.5e-5 where
// gelu -> 0; beyond 6 the result underflows towards 0 like the true value 1e-9). |gelu_fast - gelu| < 7e-6 absolute,
// ~300x below the bf16 rounding of the result. 12 instructions, 1 MUFU (the A&S form above needs rcp + ex2).
__device__ __forceinline__ float gelu_fast_f(float x) {
  const float ax = fabsf(x);
  float q = fmaf(2.29905825e-05f, ax, -0.000611092365f);
  q = fmaf(q, ax, 0.00719537331f);
  q = fmaf(q, ax, -0.0511841573f);
  q = fmaf(q, ax, -0.461274841f);
  q = fmaf(q, ax, -1.15017144f);
  q = fmaf(q, ax, -1.00006552f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(q));    // Phi(-|x|)
  return x * (x < 0.f ? h : 1.f - h);
}
__device__ __forceinline__ float gelu_grad_fast_f(float x) {
  float cdf, pdf;
  gelu_parts_fast(x, cdf, pdf);
  return fmaf(x, pdf, cdf);
}
// dtype-dispatched: exact for fp32 activations, fast for bf16 activations
template <typename T> __device__ __forceinline__ float gelu_t(float x) { return gelu_f(x); }
template <> __device__ __forceinline__ float gelu_t<bf16>(float x) { return gelu_fast_f(x); }
template <typename T> __device__ __forceinline__ float gelu_grad_t(float x) { return gelu_grad_f(x); }
template <> __device__ __forceinline__ float gelu_grad_t<bf16>(float x) { return gelu_grad_fast_f(x); }

// ----------------------------------------------------------------------------------------------
// warp / block reductions
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` is a __shared__ float[32]. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// Cheapest variant, for kernels that own a whole dropout "stream" and walk it in 32-element chunks (the fused attention
// kernels: one stream per (batch, head); a chunk = 32 consecutive keys of one query row). One hash per CHUNK seeds a
// 32-step LCG whose t-th state is reached directly with the jump-ahead constants (x_t = x_0 * A^t + C_t, one IMAD per
// element), and element t is kept iff x_t >= rate * 2^32. Any kernel can regenerate element (row, chunk, t) on its own.
struct DropKey { uint32_t k1, k2m, thr; };
__host__ __device__ __forceinline__ DropKey make_drop_key(uint64_t seed, uint64_t stream, uint32_t thr) {
  uint64_t z = seed + (stream + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  DropKey k;
  k.k1 = (uint32_t)z;
  k.k2m = (uint32_t)(z >> 32) * 0x846ca68bu;
  k.thr = thr;
  return k;
}
// chunk seed: a full-avalanche hash of the chunk index (row * chunks_per_row + chunk)
__host__ __device__ __forceinline__ uint32_t drop_chunk_seed(const DropKey& k, uint32_t chunk_index) {
  uint32_t x = chunk_index ^ k.k1;
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x = x * 0x846ca68bu + k.k2m;
  x ^= x >> 16;
  return x;
}
__host__ __device__ constexpr uint32_t lcg_a(int t) { return t == 0 ? 1u : lcg_a(t - 1) * 1664525u; }
__host__ __device__ constexpr uint32_t lcg_c(int t) { return t == 0 ? 0u : lcg_c(t - 1) * 1664525u + 1013904223u; }
// element t (0..31) of the chunk seeded with x0 (t + 1 LCG steps from the seed, so the seed itself is never used)
template <int T> __host__ __device__ __forceinline__ uint32_t drop_elem(uint32_t x0) {
  constexpr uint32_t A = lcg_a(T + 1), C = lcg_c(T + 1);
  return x0 * A + C;
}
__host__ __device__ __forceinline__ uint32_t drop_elem_rt(uint32_t x0, int t) {
  uint32_t x = x0;
  for (int i = 0; i <= t; ++i) x = x * 1664525u + 1013904223u;
  return x;
}

// Jump-ahead constants as a table, for kernels that reach element t of a chunk with a run-time t.
struct LcgTable { uint32_t a[32], c[32]; };
constexpr LcgTable make_lcg_table() {
  LcgTable t{};
  uint32_t a = 1u, c = 0u;
  for (int i = 0; i < 32; ++i) {
    a *= 1664525u;
    c = c * 1664525u + 1013904223u;
    t.a[i] = a; t.c[i] = c;
  }
  return t;
}
static __constant__ LcgTable kLcg = make_lcg_table();

// ---- dropout over a flat tensor (GEMM epilogues, element-wise passes, embeddings): the same chunked generator -------------
// Element idx of the tensor belongs to chunk idx >> 5 at position t = idx & 31; mask(idx) = [x_t >= rate * 2^32] with
// x_t = hash(seed, chunk) * A^(t+1) + C_(t+1). A kernel that owns 32 (or 8) aligned consecutive elements pays one hash per
// chunk and one IMAD + compare per element; any other kernel regenerates the identical mask element by element
// (dropout_scale), which is what keeps forward and backward masks in agreement across kernels.
__device__ __forceinline__ DropKey flat_drop_key(uint64_t seed, uint32_t thr) { return make_drop_key(seed, 0, thr); }
// returns the multiplier to apply: 0 (dropped) or 1/(1-rate) (kept). thr = rate * 2^32.
__device__ __forceinline__ float dropout_scale(const DropKey& k, uint64_t idx, float inv_keep) {
  const uint32_t x0 = drop_chunk_seed(k, (uint32_t)(idx >> 5));
  const int t = (int)(idx & 31);
  return (x0 * kLcg.a[t] + kLcg.c[t]) >= k.thr ? inv_keep : 0.f;
}
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, uint32_t thr, float inv_keep) {
  return dropout_scale(flat_drop_key(seed, thr), idx, inv_keep);
}
// multipliers of the 8 consecutive elements idx0 .. idx0+7, idx0 a multiple of 8 (they share a chunk)
__device__ __forceinline__ void dropout_scale8(const DropKey& k, uint64_t idx0, float inv_keep, float (&s)[8]) {
  const uint32_t x0 = drop_chunk_seed(k, (uint32_t)(idx0 >> 5));
  const int t0 = (int)(idx0 & 31);
  uint32_t x = x0 * kLcg.a[t0] + kLcg.c[t0];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] = x >= k.thr ? inv_keep : 0.f;
    x = x * 1664525u + 1013904223u;
  }
}
// v[0..31] *= multipliers of the chunk `chunk` (a thread that owns a whole aligned chunk); compile-time jump-ahead
template <int T> struct FlatDropUnroll {
  static __device__ __forceinline__ void apply(float* v, uint32_t x0, uint32_t thr, float inv_keep) {
    v[T] = drop_elem<T>(x0) >= thr ? v[T] * inv_keep : 0.f;
    FlatDropUnroll<T + 1>::apply(v, x0, thr, inv_keep);
  }
};
template <> struct FlatDropUnroll<32> {
  static __device__ __forceinline__ void apply(float*, uint32_t, uint32_t, float) {}
};
__device__ __forceinline__ void dropout_apply_chunk(float* v, const DropKey& k, uint32_t chunk, float inv_keep) {
  FlatDropUnroll<0>::apply(v, drop_chunk_seed(k, chunk), k.thr, inv_keep);
}

// every dropout kernel folds the device-resident salt (Ctx::d_state[0], 0 unless ts_step_state_* is used) into its seed
__device__ __forceinline__ unsigned long long salted_seed(unsigned long long seed, const unsigned long long* salt) {
  return seed + __ldg(salt) * 0x9E3779B97F4A7C15ull;
}

}  // namespace ts
