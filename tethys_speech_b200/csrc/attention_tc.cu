// K10 — fused attention forward / backward on tcgen05 (bf16 operands, fp32 accumulate in TMEM), head_dim 64.
//
// Replaces, per attention site, the reference's  q k^T -> (+mask) -> softmax -> dropout -> . v  chain and its autodiff
// transpose (W:147-167 MultiHeadAttention.call; V:348-362 Wav2Vec2MultiHeadAttention.call): the [B,H,Tq,Tk] score /
// probability tensors never reach HBM. Three kernels, all with the same warp roles as the GEMM engine
// (warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer, warps 2-5 = one thread per tile row):
//
//   attn_fwd_kernel     CTA = (128 query rows, head, batch); loops over 128-row K/V tiles:
//                         S = Q K^T (TMEM) -> online softmax in registers (lazy rescale) -> P (bf16, smem) -> O += P V (TMEM)
//                       writes O (bf16) and the row statistics (running max m, log of the row sum) for backward.
//   attn_bwd_dq_kernel  CTA = (128 query rows, head, batch); recomputes S and dP = dO V^T per K/V tile,
//                         dS = P o (dP o Z - D) * scale -> smem -> dQ += dS K (TMEM); also produces D = rowsum(dO o O).
//   attn_bwd_dkv_kernel CTA = (128 key rows, head, batch); loops over query tiles: S, dP as above,
//                         dV += (P o Z)^T dO, dK += dS^T Q (P / dS are read from smem as MN-major A operands).
// Z is the dropout keep-mask / keep-probability, regenerated from (seed, element index) in all three kernels.
//
// mask_mode 1 is the reference's decoder mask (W:150-154, W:416-418): -1e9 is ADDED in fp32 to the scores of keys
// j <= i ("anti-causal", bug-compatible), literally, so a fully masked row degenerates to the same uniform
// distribution as in TensorFlow (fp32 absorption). The row statistics keep the max and the log-sum
// separate for the same reason (a single fp32 logsumexp near -1e9 has an ulp of 64).
#include <math.h>
#include "common.cuh"
#include "ops.cuh"
#include "ptx.cuh"

namespace ts {

int get_tmap(Ctx* ctx, CUtensorMap* out, const void* base, const uint64_t d[4], const uint64_t sbytes[3], uint32_t box0,
             uint32_t box1, bool f32);

namespace {

constexpr int AT_M = 128;  // query rows per tile
constexpr int AT_N = 128;  // key/value rows per tile
constexpr int AT_D = 64;   // head dim
constexpr int kTile = AT_M * AT_D * 2;  // 16 KB: one [128 x 64] bf16 operand tile (128-byte rows, SWIZZLE_128B)
constexpr int kPTile = AT_M * AT_N * 2; // 32 KB: P / dS tile = two 64-column atoms of 16 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThr = 5.5f;     // lazy rescale: keep a stale row max while the true one is < e^5.5 larger

struct AttnParams {
  int B, nh, Tq, Tk;
  float scale;
  uint32_t drop_thr; float inv_keep; unsigned long long seed;
  const unsigned long long* salt;   // device-resident dropout salt (Ctx::d_state)
  int drop_pitch;   // 32-key chunks per query row in the per-(batch, head) dropout stream (chunk index = row * pitch + key / 32)
  bf16* o; bf16* o_lo; long long o_ld, o_bs;
  float* stats;   // [B, nh, Tq, 2]
  const bf16* o_in; const bf16* d_o;  // same layout as o
  float* dsum;    // [B, nh, Tq]
  bf16* dq; long long dq_ld, dq_bs;
  bf16 *dk, *dv; long long dkv_ld, dkv_bs;
};

template <int MASK>
__device__ __forceinline__ float score_of(float acc, float scale, int i, int j, int Tk) {
  float s = acc * scale;
  if (MASK == 1 && j <= i) s += -1e9f;  // literal fp32 add (absorption is part of the reference's semantics)
  return j < Tk ? s : -INFINITY;
}

// address of the 16-byte piece holding columns [8*piece, 8*piece+8) of row r in a [128 x 128] bf16 tile stored as two
// 64-column SWIZZLE_128B atoms (the layout tcgen05 reads as a K-major A operand, or as an MN-major one transposed)
__device__ __forceinline__ uint32_t p_tile_addr(uint32_t base, int r, int piece) {
  return base + (piece >> 3) * (kTile) + r * 128 + (((piece & 7) ^ (r & 7)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// K-major [128 x 64] operand (Q, K as B of S, dO, V as B of dP): k-step kk covers 16 columns = 32 bytes
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int kk) { return ptx::make_smem_desc(tile + kk * 32, 16, 1024); }
// P / dS as K-major A [128 rows x 128 k]: 8 k-steps over two atoms
__device__ __forceinline__ uint64_t desc_p_kmajor(uint32_t tile, int kk) {
  return ptx::make_smem_desc(tile + (kk >> 2) * kTile + (kk & 3) * 32, 16, 1024);
}
// [128 k-rows x 64] tile read as an MN-major B operand (V in P.V, K in dS.K, dO / Q in the dV / dK products)
__device__ __forceinline__ uint64_t desc_mn_b(uint32_t tile, int kk) { return ptx::make_smem_desc(tile + kk * 2048, 8192, 1024); }
#define TS_TRY_RC(expr)   \
  do {                   \
    int _rc = (expr);    \
    if (_rc) return _rc; \
  } while (0)

// the kernels carve SWIZZLE_128B tiles straight from the dynamic smem window: it must start 1024-byte aligned
__device__ __forceinline__ bool smem_aligned(const void* smem, int* watchdog) {
  if ((ptx::smem_u32(smem) & 1023u) == 0) return true;
  if (threadIdx.x == 0 && watchdog) atomicExch(watchdog, 99);
  return false;
}

// zero the dropped ones among a thread's 32 consecutive chunk elements (compile-time unrolled LCG jump-ahead)
template <int T> struct DropUnroll {
  static __device__ __forceinline__ void apply(float (&v)[32], uint32_t x0, uint32_t thr) {
    v[T] = drop_elem<T>(x0) >= thr ? v[T] : 0.f;
    DropUnroll<T + 1>::apply(v, x0, thr);
  }
  // z[t] = keep ? inv_keep : 0
  static __device__ __forceinline__ void scales(float (&z)[32], uint32_t x0, uint32_t thr, float inv_keep) {
    z[T] = drop_elem<T>(x0) >= thr ? inv_keep : 0.f;
    DropUnroll<T + 1>::scales(z, x0, thr, inv_keep);
  }
};
template <> struct DropUnroll<32> {
  static __device__ __forceinline__ void apply(float (&)[32], uint32_t, uint32_t) {}
  static __device__ __forceinline__ void scales(float (&)[32], uint32_t, uint32_t, float) {}
};
__device__ __forceinline__ void drop_apply32(float (&v)[32], uint32_t x0, uint32_t thr) { DropUnroll<0>::apply(v, x0, thr); }

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
// 16 softmax warps per CTA: warp % 4 = TMEM lane quarter (32 rows), (warp - 2) / 4 = 32-column chunk of the score tile.
// S is double buffered in TMEM and P in smem, K/V run through a 3-stage TMA ring: the tensor pipe computes S(j+1)
// and P.V(j) while the softmax warps work on tile j+1.
constexpr int kFwdEw = 16;
constexpr int kFwdThreads = 64 + kFwdEw * 32;
constexpr int kFwdStages = 3;
constexpr int kFwdSmem = kTile /*Q*/ + 2 * kFwdStages * kTile /*K, V*/ + 2 * kPTile /*P[2]*/ + 2 * 4 * 128 * 4 /*row-max exchange*/ +
                         4 * 128 * 4 /*row-sum exchange*/ + 256 /*barriers*/;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void quarter_sync(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }

template <int MASK>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const AttnParams p, int* watchdog) {
  ts::pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;  // 1024-byte aligned by declaration (no static shared memory in this kernel)
  if (!smem_aligned(smem, watchdog)) return;
  const uint32_t sQ = ptx::smem_u32(smem), sK = sQ + kTile, sV = sK + kFwdStages * kTile, sP = sV + kFwdStages * kTile;
  float* smax = reinterpret_cast<float*>(smem + kTile * (1 + 2 * kFwdStages) + 2 * kPTile);  // [2][4][128]
  float* ssum = smax + 2 * 4 * 128;                                                          // [4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ssum + 4 * 128);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // 3
  uint64_t* kv_empty = bars + 4;    // 3
  uint64_t* s_full = bars + 7;      // 2
  uint64_t* p_full = bars + 9;      // 2 (kFwdEw warp arrivals each)
  uint64_t* pv_done = bars + 11;    // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_M, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Tk + AT_N - 1) / AT_N;

  if (threadIdx.x == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < kFwdStages; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&s_full[s], 1); ptx::mbar_init(&p_full[s], kFwdEw); ptx::mbar_init(&pv_done[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 256;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_q); ptx::prefetch_tmap(&tm_k); ptx::prefetch_tmap(&tm_v);
      ptx::mbar_expect_tx(q_full, kTile);
      ptx::tma_load_4d(sQ, &tm_q, q_full, 0, q0, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j % kFwdStages;
        if (!ptx::mbar_wait(&kv_empty[s], ((j / kFwdStages) & 1) ^ 1, watchdog, 11)) break;
        ptx::mbar_expect_tx(&kv_full[s], 2 * kTile);
        ptx::tma_load_4d(sK + s * kTile, &tm_k, &kv_full[s], 0, j * AT_N, h, b);
        ptx::tma_load_4d(sV + s * kTile, &tm_v, &kv_full[s], 0, j * AT_N, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(AT_M, AT_N, 0, 0);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(AT_M, AT_D, 0, 1);
      auto issue_s = [&](int j) {
        const int s = j % kFwdStages;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tS + (j & 1) * AT_N, desc_kmajor(sQ, kk), desc_kmajor(sK + s * kTile, kk), idesc_s, kk > 0 ? 1u : 0u);
        ptx::umma_commit(&s_full[j & 1]);
      };
      bool ok = ptx::mbar_wait(q_full, 0, watchdog, 12) && ptx::mbar_wait(&kv_full[0], 0, watchdog, 13);
      if (ok) { ptx::tc_fence_after(); issue_s(0); }
      for (int j = 0; j < nkv && ok; ++j) {
        if (j + 1 < nkv) {  // S of the next tile goes first: its buffer was released by p_full(j-1), waited last iteration
          if (!ptx::mbar_wait(&kv_full[(j + 1) % kFwdStages], ((j + 1) / kFwdStages) & 1, watchdog, 13)) break;
          ptx::tc_fence_after();
          issue_s(j + 1);
        }
        if (!ptx::mbar_wait(&p_full[j & 1], (j >> 1) & 1, watchdog, 14)) break;
        ptx::tc_fence_after();
        const int s = j % kFwdStages;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          ptx::umma_f16(tO, desc_p_kmajor(sP + (j & 1) * kPTile, kk), desc_mn_b(sV + s * kTile, kk), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        ptx::umma_commit(&kv_empty[s]);
        ptx::umma_commit(&pv_done[j & 1]);
      }
    }
  } else {
    const int q = warp & 3, chunk = (warp - 2) >> 2;
    const int r = q * 32 + lane;             // row inside the tile
    const int i = q0 + r;                    // query index
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const DropKey dkey = make_drop_key(p.drop_thr ? salted_seed(p.seed, p.salt) : p.seed, (unsigned long long)(b * p.nh + h), p.drop_thr);
    const uint32_t drow = (uint32_t)i * (uint32_t)p.drop_pitch;   // chunk index of this row's first 32-key chunk in the stream
    const float c1 = p.scale * kLog2e;
    float m_run = -INFINITY, l_part = 0.f;
    bool ok = true;
    for (int j = 0; j < nkv && ok; ++j) {
      const int bsel = j & 1;
      if (!ptx::mbar_wait(&s_full[bsel], (j >> 1) & 1, watchdog, 15)) { ok = false; break; }
      ptx::tc_fence_after();
      const int col0 = j * AT_N + chunk * 32;
      uint32_t rg[32];
      ptx::tmem_ld_32x32(tS + bsel * AT_N + lane_off + chunk * 32, rg);
      ptx::tmem_ld_wait();
      float sv[32];
      float mx = -INFINITY;
      const bool ragged = col0 + 32 > p.Tk;
      if (MASK == 0 && !ragged) {
#pragma unroll
        for (int t = 0; t < 32; ++t) { sv[t] = __uint_as_float(rg[t]); mx = fmaxf(mx, sv[t]); }
        mx *= p.scale;   // scale > 0
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) { sv[t] = score_of<MASK>(__uint_as_float(rg[t]), p.scale, i, col0 + t, p.Tk); mx = fmaxf(mx, sv[t]); }
      }
      // row max over the four column chunks (one warp each)
      smax[(bsel * 4 + chunk) * 128 + r] = mx;
      quarter_sync(q);
      const float* sm = smax + bsel * 4 * 128 + r;
      const float mrow = fmaxf(fmaxf(sm[0], sm[128]), fmaxf(sm[256], sm[384]));
      const bool need = __any_sync(0xffffffffu, mrow > m_run + kRescaleThr);
      if (need) {
        const float m_new = fmaxf(m_run, mrow);
        const float alpha = ex2f((m_run - m_new) * kLog2e);  // 0 on the first tile
        l_part *= alpha;
        if (j > 0) {  // rescale this thread's 16 columns of O once P.V of the previous tile has retired
          if (!ptx::mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1, watchdog, 16)) { ok = false; break; }
          ptx::tc_fence_after();
          uint32_t ro[16];
          ptx::tmem_ld_32x16(tO + lane_off + chunk * 16, ro);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 16; ++t) ro[t] = __float_as_uint(__uint_as_float(ro[t]) * alpha);
          ptx::tmem_st_32x16(tO + lane_off + chunk * 16, ro);
          ptx::tmem_st_wait();
        }
        m_run = m_new;
      }
      if (MASK == 0 && !ragged) {
        const float nm = -m_run * kLog2e;
#pragma unroll
        for (int t = 0; t < 32; ++t) { sv[t] = ex2f(fmaf(sv[t], c1, nm)); l_part += sv[t]; }
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) { sv[t] = ex2f((sv[t] - m_run) * kLog2e); l_part += sv[t]; }
      }
      if (p.drop_thr) {  // dropped probabilities become 0; the common factor 1/keep is applied to O at the end
        const uint32_t x0 = drop_chunk_seed(dkey, drow + ((uint32_t)col0 >> 5));
        drop_apply32(sv, x0, dkey.thr);
      }
      if (j >= 2) {  // P buffer `bsel` is free once P.V of tile j-2 has retired
        if (!ptx::mbar_wait(&pv_done[bsel], ((j - 2) >> 1) & 1, watchdog, 18)) { ok = false; break; }
      }
      const uint32_t pt = sP + bsel * kPTile;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        sts128(p_tile_addr(pt, r, chunk * 4 + t), pack2(sv[8 * t], sv[8 * t + 1]), pack2(sv[8 * t + 2], sv[8 * t + 3]),
               pack2(sv[8 * t + 4], sv[8 * t + 5]), pack2(sv[8 * t + 6], sv[8 * t + 7]));
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[bsel]);
    }
    // row sum over the four chunks, then each thread normalises and stores 16 of the 64 output columns
    ssum[chunk * 128 + r] = l_part;
    quarter_sync(q);
    const float l_row = ssum[r] + ssum[128 + r] + ssum[256 + r] + ssum[384 + r];
    if (ok && ptx::mbar_wait(&pv_done[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1, watchdog, 17)) {
      ptx::tc_fence_after();
      uint32_t ro[16];
      ptx::tmem_ld_32x16(tO + lane_off + chunk * 16, ro);
      ptx::tmem_ld_wait();
      if (i < p.Tq) {
        const float inv = p.inv_keep / l_row;
        const long long ooff = (long long)b * p.o_bs + (long long)i * p.o_ld + h * AT_D + chunk * 16;
        float ov[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) ov[t] = __uint_as_float(ro[t]) * inv;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          uint4 u;
          u.x = pack2(ov[8 * t], ov[8 * t + 1]); u.y = pack2(ov[8 * t + 2], ov[8 * t + 3]);
          u.z = pack2(ov[8 * t + 4], ov[8 * t + 5]); u.w = pack2(ov[8 * t + 6], ov[8 * t + 7]);
          reinterpret_cast<uint4*>(p.o + ooff)[t] = u;
        }
        if (p.o_lo) {  // bf16 rounding residual, so that backward can form D = rowsum(dO o O) from an (almost) fp32 O
#pragma unroll
          for (int t = 0; t < 16; ++t) ov[t] -= __bfloat162float(__float2bfloat16_rn(ov[t]));
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            uint4 u;
            u.x = pack2(ov[8 * t], ov[8 * t + 1]); u.y = pack2(ov[8 * t + 2], ov[8 * t + 3]);
            u.z = pack2(ov[8 * t + 4], ov[8 * t + 5]); u.w = pack2(ov[8 * t + 6], ov[8 * t + 7]);
            reinterpret_cast<uint4*>(p.o_lo + ooff)[t] = u;
          }
        }
        if (chunk == 0) {
          float* st = p.stats + (((long long)b * p.nh + h) * p.Tq + i) * 2;
          st[0] = m_run;
          st[1] = logf(l_row);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
// Both kernels work on 128 x 64 score tiles with 8 element-wise warps (warp % 4 = TMEM lane quarter = 32 tile rows,
// (warp - 2) / 4 = which 32 of the 64 tile columns) and fit two CTAs per SM (<= 256 TMEM columns, <= 100 KB smem):
// while one CTA's element-wise warps chew on a tile, the other CTA's MMAs / TMA / barrier hand-offs run underneath.
constexpr int kBwdEw = 8;
constexpr int kBwdThreads = 64 + kBwdEw * 32;
constexpr int kHalfTile = 64 * AT_D * 2;  // 8 KB: a [64 x 64] bf16 operand tile

__device__ __forceinline__ void store_chunk32(uint32_t tile, int r, int chunk, const float (&v)[32]) {
#pragma unroll
  for (int t = 0; t < 4; ++t)
    sts128(p_tile_addr(tile, r, chunk * 4 + t), pack2(v[8 * t], v[8 * t + 1]), pack2(v[8 * t + 2], v[8 * t + 3]),
           pack2(v[8 * t + 4], v[8 * t + 5]), pack2(v[8 * t + 6], v[8 * t + 7]));
}
// 32 fp32 accumulator columns of this thread's row -> bf16 -> 64 contiguous bytes in global memory
__device__ __forceinline__ void store_row32(bf16* dst, const uint32_t (&rg)[32], float mul) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack2(__uint_as_float(rg[8 * j]) * mul, __uint_as_float(rg[8 * j + 1]) * mul);
    u.y = pack2(__uint_as_float(rg[8 * j + 2]) * mul, __uint_as_float(rg[8 * j + 3]) * mul);
    u.z = pack2(__uint_as_float(rg[8 * j + 4]) * mul, __uint_as_float(rg[8 * j + 5]) * mul);
    u.w = pack2(__uint_as_float(rg[8 * j + 6]) * mul, __uint_as_float(rg[8 * j + 7]) * mul);
    reinterpret_cast<uint4*>(dst)[j] = u;
  }
}

// ---- dQ (and D = rowsum(dO o O)): CTA = 128 query rows, loop over 64-row K/V tiles ---------------------------------
// One thread's 32 columns: p = exp(s' - m - log l) recomputed from the raw accumulators and dS' = p o (dP o Z - D)
// (the common factor `scale` is applied when dQ is stored). c1 = scale*log2(e), c0 = m*log2(e) + log2(l) (+inf: padding row).
template <int MASK, bool RAGGED>
__device__ __forceinline__ void dq_chunk(const uint32_t (&rs)[32], const uint32_t (&rd)[32], float (&ds)[32], const AttnParams& p,
                                         float c1, float c0, float m_i, float ll2, float D, int i, int col0, const DropKey& dkey,
                                         uint32_t chunk_index) {
  float z[32];
  if (p.drop_thr) DropUnroll<0>::scales(z, drop_chunk_seed(dkey, chunk_index), dkey.thr, p.inv_keep);
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    float p0;
    if (MASK == 0) {
      p0 = ex2f(fmaf(__uint_as_float(rs[t]), c1, -c0));
    } else {
      float s0 = __uint_as_float(rs[t]) * p.scale;
      if (col0 + t <= i) s0 += -1e9f;
      p0 = ex2f((s0 - m_i) * kLog2e - ll2);
    }
    if (RAGGED) {
      if (col0 + t >= p.Tk) p0 = 0.f;
    }
    if (p.drop_thr) ds[t] = p0 * fmaf(__uint_as_float(rd[t]), z[t], -D);
    else ds[t] = p0 * (__uint_as_float(rd[t]) - D);
  }
}

constexpr int kDqSmem = 2 * kTile /*Q, dO*/ + 4 * kHalfTile /*K[2], V[2]*/ + kTile /*dS [128 x 64]*/ + 128;

template <int MASK>
__global__ void __launch_bounds__(kBwdThreads, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do, const AttnParams p,
                   int* watchdog) {
  ts::pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (!smem_aligned(smem, watchdog)) return;
  const uint32_t sQ = ptx::smem_u32(smem), sDO = sQ + kTile, sK = sDO + kTile, sV = sK + 2 * kHalfTile, sDS = sV + 2 * kHalfTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTile + 4 * kHalfTile);
  uint64_t* q_full = bars;        // Q + dO
  uint64_t* kv_full = bars + 1;   // 2
  uint64_t* kv_empty = bars + 3;  // 2
  uint64_t* sdp_full = bars + 5;
  uint64_t* ds_full = bars + 6;   // kBwdEw warp arrivals
  uint64_t* dq_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_M, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Tk + 63) / 64;
  if (threadIdx.x == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    ptx::mbar_init(sdp_full, 1);
    ptx::mbar_init(ds_full, kBwdEw);
    ptx::mbar_init(dq_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + 64, tDQ = tmem + 128;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_q); ptx::prefetch_tmap(&tm_k); ptx::prefetch_tmap(&tm_v); ptx::prefetch_tmap(&tm_do);
      ptx::mbar_expect_tx(q_full, 2 * kTile);
      ptx::tma_load_4d(sQ, &tm_q, q_full, 0, q0, h, b);
      ptx::tma_load_4d(sDO, &tm_do, q_full, 0, q0, h, b);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        if (!ptx::mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1, watchdog, 21)) break;
        ptx::mbar_expect_tx(&kv_full[s], 2 * kHalfTile);
        ptx::tma_load_4d(sK + s * kHalfTile, &tm_k, &kv_full[s], 0, j * 64, h, b);
        ptx::tma_load_4d(sV + s * kHalfTile, &tm_v, &kv_full[s], 0, j * 64, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(AT_M, 64, 0, 0);
      constexpr uint32_t idesc_dq = ptx::make_idesc_bf16(AT_M, AT_D, 0, 1);
      auto issue_sdp = [&](int s) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tS, desc_kmajor(sQ, kk), desc_kmajor(sK + s * kHalfTile, kk), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tDP, desc_kmajor(sDO, kk), desc_kmajor(sV + s * kHalfTile, kk), idesc_s, kk > 0 ? 1u : 0u);
        ptx::umma_commit(sdp_full);
      };
      bool ok = ptx::mbar_wait(q_full, 0, watchdog, 22) && ptx::mbar_wait(&kv_full[0], 0, watchdog, 23);
      if (ok) { ptx::tc_fence_after(); issue_sdp(0); }
      for (int j = 0; j < nkv && ok; ++j) {
        const int s = j & 1;
        if (!ptx::mbar_wait(ds_full, j & 1, watchdog, 24)) break;
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tDQ, desc_kmajor(sDS, kk), desc_mn_b(sK + s * kHalfTile, kk), idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
        ptx::umma_commit(&kv_empty[s]);
        ptx::umma_commit(dq_done);
        if (j + 1 < nkv) {
          if (!ptx::mbar_wait(&kv_full[s ^ 1], ((j + 1) >> 1) & 1, watchdog, 23)) break;
          ptx::tc_fence_after();
          issue_sdp(s ^ 1);
        }
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int i = q0 + r;
    const bool live = i < p.Tq;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const DropKey dkey = make_drop_key(p.drop_thr ? salted_seed(p.seed, p.salt) : p.seed, (unsigned long long)(b * p.nh + h), p.drop_thr);
    const uint32_t drow = (uint32_t)i * (uint32_t)p.drop_pitch;
    const long long srow = ((long long)b * p.nh + h) * p.Tq + i;
    float m_i = 0.f, ll2 = INFINITY, D = 0.f;
    if (live) {
      m_i = p.stats[srow * 2];
      ll2 = p.stats[srow * 2 + 1] * kLog2e;
      const uint4* po = reinterpret_cast<const uint4*>(p.o_in + (long long)b * p.o_bs + (long long)i * p.o_ld + h * AT_D);
      const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + (long long)b * p.o_bs + (long long)i * p.o_ld + h * AT_D);
      const uint4* pl = p.o_lo ? reinterpret_cast<const uint4*>(p.o_lo + (long long)b * p.o_bs + (long long)i * p.o_ld + h * AT_D) : nullptr;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const uint4 a = __ldg(po + t), g = __ldg(pd + t);
        const uint4 lo = pl ? __ldg(pl + t) : make_uint4(0, 0, 0, 0);
        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&g);
        const __nv_bfloat162* lh = reinterpret_cast<const __nv_bfloat162*>(&lo);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = __bfloat1622float2(ah[e]), y = __bfloat1622float2(gh[e]), z = __bfloat1622float2(lh[e]);
          D = fmaf(x.x + z.x, y.x, D);
          D = fmaf(x.y + z.y, y.y, D);
        }
      }
      if (half == 0) p.dsum[srow] = D;
    }
    const float c1 = p.scale * kLog2e, c0 = m_i * kLog2e + ll2;
    bool ok = true;
    for (int j = 0; j < nkv && ok; ++j) {
      if (!ptx::mbar_wait(sdp_full, j & 1, watchdog, 25)) { ok = false; break; }
      ptx::tc_fence_after();
      const int col0 = j * 64 + half * 32;
      uint32_t rs[32], rd[32];
      ptx::tmem_ld_32x32(tS + lane_off + half * 32, rs);
      ptx::tmem_ld_32x32(tDP + lane_off + half * 32, rd);
      ptx::tmem_ld_wait();
      float ds[32];
      const uint32_t cidx = drow + ((uint32_t)col0 >> 5);
      if (col0 + 32 > p.Tk) dq_chunk<MASK, true>(rs, rd, ds, p, c1, c0, m_i, ll2, D, i, col0, dkey, cidx);
      else dq_chunk<MASK, false>(rs, rd, ds, p, c1, c0, m_i, ll2, D, i, col0, dkey, cidx);
      if (j > 0 && !ptx::mbar_wait(dq_done, (j - 1) & 1, watchdog, 26)) { ok = false; break; }  // dS smem free again
      store_chunk32(sDS, r, half, ds);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ds_full);
    }
    if (ok && ptx::mbar_wait(dq_done, (nkv - 1) & 1, watchdog, 27)) {
      ptx::tc_fence_after();
      uint32_t rg[32];
      ptx::tmem_ld_32x32(tDQ + lane_off + half * 32, rg);
      ptx::tmem_ld_wait();
      if (live) store_row32(p.dq + (long long)b * p.dq_bs + (long long)i * p.dq_ld + h * AT_D + half * 32, rg, p.scale);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 256);
}

__device__ __forceinline__ uint32_t lcg_a_rt(int t) { uint32_t a = 1u; for (int i = 0; i < t; ++i) a *= 1664525u; return a; }
__device__ __forceinline__ uint32_t lcg_c_rt(int t) { uint32_t c = 0u; for (int i = 0; i < t; ++i) c = c * 1664525u + 1013904223u; return c; }

// ---- dK, dV: CTA = 128 key rows, loop over 64-row query tiles, TRANSPOSED scores ------------------------------------
// S^T = K Q^T and dP^T = V dO^T put the key index on the TMEM lanes, so a thread owns one key row and 32 query columns,
// P^T / dS^T land in smem as plain K-major A operands of dV += P^T dO and dK += dS^T Q, and the accumulators need no
// transposes. Per-query statistics (c0 = m*log2e + log2 l, D) are staged in smem per tile and read as broadcasts.
constexpr int kDkvSmem = 2 * kTile /*K, V*/ + 4 * kHalfTile /*Q[2], dO[2]*/ + 2 * kTile /*P^T, dS^T*/ + 2 * 4 * 64 * 4 /*stats*/ + 128;

template <int MASK>
__global__ void __launch_bounds__(kBwdThreads, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do, const AttnParams p,
                    int* watchdog) {
  ts::pdl_enter();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (!smem_aligned(smem, watchdog)) return;
  const uint32_t sK = ptx::smem_u32(smem), sV = sK + kTile, sQ = sV + kTile, sDO = sQ + 2 * kHalfTile, sPT = sDO + 2 * kHalfTile,
                 sDST = sPT + kTile;
  float* s_stat = reinterpret_cast<float*>(smem + 4 * kTile + 4 * kHalfTile);  // [2 buffers][4: c0, D, m, ll2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 2 * 4 * 64);
  uint64_t* kv_full = bars;        // K + V
  uint64_t* q_full = bars + 1;     // 2 (Q + dO)
  uint64_t* q_empty = bars + 3;    // 2
  uint64_t* sdp_full = bars + 5;
  uint64_t* pds_full = bars + 6;   // kBwdEw warp arrivals
  uint64_t* acc_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * AT_N, h = blockIdx.y, b = blockIdx.z;
  const int nq = (p.Tq + 63) / 64;
  if (threadIdx.x == 0) {
    ptx::mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&q_full[s], 1); ptx::mbar_init(&q_empty[s], 1); }
    ptx::mbar_init(sdp_full, 1);
    ptx::mbar_init(pds_full, kBwdEw);
    ptx::mbar_init(acc_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tST = tmem, tDPT = tmem + 64, tDV = tmem + 128, tDK = tmem + 192;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_q); ptx::prefetch_tmap(&tm_k); ptx::prefetch_tmap(&tm_v); ptx::prefetch_tmap(&tm_do);
      ptx::mbar_expect_tx(kv_full, 2 * kTile);
      ptx::tma_load_4d(sK, &tm_k, kv_full, 0, kv0, h, b);
      ptx::tma_load_4d(sV, &tm_v, kv_full, 0, kv0, h, b);
      for (int it = 0; it < nq; ++it) {
        const int s = it & 1;
        if (!ptx::mbar_wait(&q_empty[s], ((it >> 1) & 1) ^ 1, watchdog, 31)) break;
        ptx::mbar_expect_tx(&q_full[s], 2 * kHalfTile);
        ptx::tma_load_4d(sQ + s * kHalfTile, &tm_q, &q_full[s], 0, it * 64, h, b);
        ptx::tma_load_4d(sDO + s * kHalfTile, &tm_do, &q_full[s], 0, it * 64, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(AT_N, 64, 0, 0);      // [128 keys] x [64 queries]
      constexpr uint32_t idesc_acc = ptx::make_idesc_bf16(AT_N, AT_D, 0, 1);  // A = P^T / dS^T (K-major), B = dO / Q (MN-major)
      auto issue_sdp = [&](int s) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tST, desc_kmajor(sK, kk), desc_kmajor(sQ + s * kHalfTile, kk), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tDPT, desc_kmajor(sV, kk), desc_kmajor(sDO + s * kHalfTile, kk), idesc_s, kk > 0 ? 1u : 0u);
        ptx::umma_commit(sdp_full);
      };
      bool ok = ptx::mbar_wait(kv_full, 0, watchdog, 32) && ptx::mbar_wait(&q_full[0], 0, watchdog, 33);
      if (ok) { ptx::tc_fence_after(); issue_sdp(0); }
      for (int it = 0; it < nq && ok; ++it) {
        const int s = it & 1;
        if (!ptx::mbar_wait(pds_full, it & 1, watchdog, 34)) break;
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tDV, desc_kmajor(sPT, kk), desc_mn_b(sDO + s * kHalfTile, kk), idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_f16(tDK, desc_kmajor(sDST, kk), desc_mn_b(sQ + s * kHalfTile, kk), idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
        ptx::umma_commit(&q_empty[s]);
        ptx::umma_commit(acc_done);
        if (it + 1 < nq) {
          if (!ptx::mbar_wait(&q_full[s ^ 1], ((it + 1) >> 1) & 1, watchdog, 33)) break;
          ptx::tc_fence_after();
          issue_sdp(s ^ 1);
        }
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int jrow = kv0 + r;                 // key index of this thread
    const bool rowlive = jrow < p.Tk;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int et = (warp - 2) * 32 + lane;    // 0..255 among the element-wise threads
    const float c1 = p.scale * kLog2e;
    const DropKey dkey = make_drop_key(p.drop_thr ? salted_seed(p.seed, p.salt) : p.seed, (unsigned long long)(b * p.nh + h), p.drop_thr);
    // dropout: this warp's 32 key rows are ONE chunk (index jrow >> 5) of every query row; lane t hashes the chunk seed of
    // query column t, the seed of column c is fetched by shuffle and advanced to this lane's position with its own
    // jump-ahead constants (a_l, c_l)
    const uint32_t my_chunk = (uint32_t)jrow >> 5;
    const uint32_t a_l = lcg_a_rt(lane + 1), c_l = lcg_c_rt(lane + 1);
    bool ok = true;
    for (int it = 0; it < nq && ok; ++it) {
      const int buf = it & 1;
      float* st = s_stat + buf * 4 * 64;
      if (et < 64) {  // stage the per-query statistics of this tile
        const int i = it * 64 + et;
        float m_i = 0.f, ll2 = INFINITY, D = 0.f;
        if (i < p.Tq) {
          const long long srow = ((long long)b * p.nh + h) * p.Tq + i;
          m_i = p.stats[srow * 2]; ll2 = p.stats[srow * 2 + 1] * kLog2e; D = p.dsum[srow];
        }
        st[et] = m_i * kLog2e + ll2; st[64 + et] = D; st[128 + et] = m_i; st[192 + et] = ll2;
      }
      asm volatile("bar.sync 5, 256;" ::: "memory");
      if (!ptx::mbar_wait(sdp_full, it & 1, watchdog, 35)) { ok = false; break; }
      ptx::tc_fence_after();
      uint32_t rs[32], rd[32];
      ptx::tmem_ld_32x32(tST + lane_off + half * 32, rs);
      ptx::tmem_ld_32x32(tDPT + lane_off + half * 32, rd);
      ptx::tmem_ld_wait();
      float ds[32], pz[32];
      const int i0 = it * 64 + half * 32;     // first query column of this thread
      const float* sc0 = st + half * 32;
      const float* sD = st + 64 + half * 32;
      // chunk seed of query column (i0 + lane) for this warp's key chunk
      const uint32_t xseed = p.drop_thr ? drop_chunk_seed(dkey, (uint32_t)(i0 + lane) * (uint32_t)p.drop_pitch + my_chunk) : 0u;
#pragma unroll
      for (int t = 0; t < 32; t += 2) {
        float p0, p1;
        if (MASK == 0) {
          p0 = ex2f(fmaf(__uint_as_float(rs[t]), c1, -sc0[t]));
          p1 = ex2f(fmaf(__uint_as_float(rs[t + 1]), c1, -sc0[t + 1]));
        } else {
          float s0 = __uint_as_float(rs[t]) * p.scale, s1 = __uint_as_float(rs[t + 1]) * p.scale;
          if (jrow <= i0 + t) s0 += -1e9f;
          if (jrow <= i0 + t + 1) s1 += -1e9f;
          p0 = ex2f((s0 - st[128 + half * 32 + t]) * kLog2e - st[192 + half * 32 + t]);
          p1 = ex2f((s1 - st[128 + half * 32 + t + 1]) * kLog2e - st[192 + half * 32 + t + 1]);
        }
        if (!rowlive) { p0 = 0.f; p1 = 0.f; }
        if (p.drop_thr) {
          const uint32_t w0 = __shfl_sync(0xffffffffu, xseed, t) * a_l + c_l;
          const uint32_t w1 = __shfl_sync(0xffffffffu, xseed, t + 1) * a_l + c_l;
          const bool k0 = w0 >= dkey.thr, k1 = w1 >= dkey.thr;
          ds[t] = p0 * fmaf(__uint_as_float(rd[t]), k0 ? p.inv_keep : 0.f, -sD[t]);
          ds[t + 1] = p1 * fmaf(__uint_as_float(rd[t + 1]), k1 ? p.inv_keep : 0.f, -sD[t + 1]);
          pz[t] = k0 ? p0 : 0.f;
          pz[t + 1] = k1 ? p1 : 0.f;
        } else {
          ds[t] = p0 * (__uint_as_float(rd[t]) - sD[t]);
          ds[t + 1] = p1 * (__uint_as_float(rd[t + 1]) - sD[t + 1]);
          pz[t] = p0; pz[t + 1] = p1;
        }
      }
      if (it > 0 && !ptx::mbar_wait(acc_done, (it - 1) & 1, watchdog, 36)) { ok = false; break; }  // P^T / dS^T smem free again
      store_chunk32(sPT, r, half, pz);
      store_chunk32(sDST, r, half, ds);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(pds_full);
    }
    if (ok && ptx::mbar_wait(acc_done, (nq - 1) & 1, watchdog, 37)) {
      ptx::tc_fence_after();
      uint32_t rg[32];
      ptx::tmem_ld_32x32(tDV + lane_off + half * 32, rg);
      ptx::tmem_ld_wait();
      if (rowlive) store_row32(p.dv + (long long)b * p.dkv_bs + (long long)jrow * p.dkv_ld + h * AT_D + half * 32, rg, p.inv_keep);
      ptx::tmem_ld_32x32(tDK + lane_off + half * 32, rg);
      ptx::tmem_ld_wait();
      if (rowlive) store_row32(p.dk + (long long)b * p.dkv_bs + (long long)jrow * p.dkv_ld + h * AT_D + half * 32, rg, p.scale);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 256);
}

static int head_tmap(Ctx* ctx, CUtensorMap* out, const void* base, long long ld, long long bs, int T, int nh, int B, int box_rows = 128) {
  const uint64_t dims[4] = {(uint64_t)AT_D, (uint64_t)T, (uint64_t)nh, (uint64_t)B};
  const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)AT_D * 2, (uint64_t)(B > 1 ? bs : ld) * 2};
  return get_tmap(ctx, out, base, dims, str, AT_D, (uint32_t)box_rows, false);
}

static void drop_params(float drop, uint32_t* thr, float* inv_keep) {
  if (drop <= 0.f) { *thr = 0; *inv_keep = 1.f; return; }
  double t = (double)drop * 4294967296.0;
  if (t > 4294967295.0) t = 4294967295.0;
  *thr = (uint32_t)t;
  *inv_keep = 1.f / (1.f - drop);
}

}  // namespace

bool attn_tc_supported(const ts_attn_desc* d) {
  auto al = [](const void* q) { return q && (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  auto s8 = [](long long e) { return e > 0 && e % 8 == 0; };
  return d->head_dim == AT_D && d->batch > 0 && d->heads > 0 && d->tq > 0 && d->tk > 0 && al(d->q) && al(d->k) && al(d->v) &&
         al(d->o) && s8(d->q_ld) && s8(d->kv_ld) && s8(d->o_ld) && (d->batch == 1 || (s8(d->q_bs) && s8(d->kv_bs) && s8(d->o_bs))) &&
         (d->mask_mode == 0 || d->mask_mode == 1);
}

static int fill_params(Ctx* ctx, const ts_attn_desc* d, AttnParams* p) {
  TS_REQUIRE(ctx, attn_tc_supported(d), TS_EUNSUPPORTED,
             "attention: needs head_dim 64, bf16 tensors with 16-byte aligned bases and strides that are multiples of 8");
  TS_REQUIRE(ctx, d->stats, TS_EINVAL, "attention: stats buffer [B, heads, Tq, 2] required");
  memset(p, 0, sizeof(*p));
  p->B = d->batch; p->nh = d->heads; p->Tq = d->tq; p->Tk = d->tk; p->scale = d->scale;
  drop_params(d->drop, &p->drop_thr, &p->inv_keep);
  p->seed = d->seed;
  p->salt = ctx->d_state;
  p->drop_pitch = (d->tk + 31) >> 5;
  TS_REQUIRE(ctx, (long long)d->tq * p->drop_pitch < (1ll << 32), TS_ESHAPE, "attention: Tq * Tk must stay below 2^32");
  p->o = (bf16*)d->o; p->o_lo = (bf16*)d->o_lo; p->o_ld = d->o_ld; p->o_bs = d->o_bs;
  TS_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(d->o_lo) & 15) == 0, TS_EINVAL, "attention: o_lo must be 16-byte aligned");
  p->stats = d->stats;
  return 0;
}

int attn_fwd2(Ctx* ctx, const ts_attn_desc* d, uint32_t drop_thr, float inv_keep, cudaStream_t st);   // attention_fwd2.cu

int attn_fwd(Ctx* ctx, const ts_attn_desc* d, cudaStream_t st) {
  AttnParams p;
  TS_TRY_RC(fill_params(ctx, d, &p));
  // second-generation forward (persistent, two query tiles in flight); TETHYS_ATTN_FWD=1 keeps the first one for A/B runs
  static const bool gen1 = getenv("TETHYS_ATTN_FWD") && atoi(getenv("TETHYS_ATTN_FWD")) == 1;
  if (!gen1) return attn_fwd2(ctx, d, p.drop_thr, p.inv_keep, st);
  CUtensorMap tq, tk, tv;
  TS_TRY_RC(head_tmap(ctx, &tq, d->q, d->q_ld, d->q_bs, d->tq, d->heads, d->batch));
  TS_TRY_RC(head_tmap(ctx, &tk, d->k, d->kv_ld, d->kv_bs, d->tk, d->heads, d->batch));
  TS_TRY_RC(head_tmap(ctx, &tv, d->v, d->kv_ld, d->kv_bs, d->tk, d->heads, d->batch));
  static bool attr = false;
  if (!attr) {
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    attr = true;
  }
  dim3 grid(cdiv(d->tq, AT_M), d->heads, d->batch);
  if (d->mask_mode == 0) ts::launch_k(attn_fwd_kernel<0>, grid, kFwdThreads, kFwdSmem, st, tq, tk, tv, p, ctx->d_watchdog);
  else ts::launch_k(attn_fwd_kernel<1>, grid, kFwdThreads, kFwdSmem, st, tq, tk, tv, p, ctx->d_watchdog);
  TS_LAUNCH_OK(ctx);
  return 0;
}

int attn_bwd2(Ctx* ctx, const ts_attn_desc* d, uint32_t drop_thr, float inv_keep, cudaStream_t st);   // attention_bwd2.cu

int attn_bwd(Ctx* ctx, const ts_attn_desc* d, cudaStream_t st) {
  AttnParams p;
  TS_TRY_RC(fill_params(ctx, d, &p));
  auto al = [](const void* q) { return q && (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  TS_REQUIRE(ctx, al(d->d_o) && al(d->dq) && al(d->dk) && al(d->dv) && d->dsum, TS_EINVAL, "attention backward: dO, dQ, dK, dV (16-byte aligned) and dsum required");
  TS_REQUIRE(ctx, d->dq_ld % 8 == 0 && d->dkv_ld % 8 == 0 && d->dq_bs % 8 == 0 && d->dkv_bs % 8 == 0, TS_ESHAPE, "attention backward: gradient strides must be multiples of 8");
  // fused one-kernel backward when the caller provides the fp32 dQ accumulator; TETHYS_ATTN_BWD=1 keeps the two-kernel one (A/B)
  static const bool gen1 = getenv("TETHYS_ATTN_BWD") && atoi(getenv("TETHYS_ATTN_BWD")) == 1;
  if (d->dq_accum && !gen1) {
    TS_REQUIRE(ctx, al(d->dq_accum) && al(d->o), TS_EINVAL, "attention backward: dq_accum / o must be 16-byte aligned");
    return attn_bwd2(ctx, d, p.drop_thr, p.inv_keep, st);
  }
  p.o_in = (const bf16*)d->o; p.d_o = (const bf16*)d->d_o; p.dsum = d->dsum;
  p.dq = (bf16*)d->dq; p.dq_ld = d->dq_ld; p.dq_bs = d->dq_bs;
  p.dk = (bf16*)d->dk; p.dv = (bf16*)d->dv; p.dkv_ld = d->dkv_ld; p.dkv_bs = d->dkv_bs;
  // dQ kernel: Q / dO as 128-row boxes, K / V as 64-row boxes; dK/dV kernel: the other way round
  CUtensorMap tq, tk, tv, tdo, tq64, tk64, tv64, tdo64;
  TS_TRY_RC(head_tmap(ctx, &tq, d->q, d->q_ld, d->q_bs, d->tq, d->heads, d->batch));
  TS_TRY_RC(head_tmap(ctx, &tk, d->k, d->kv_ld, d->kv_bs, d->tk, d->heads, d->batch));
  TS_TRY_RC(head_tmap(ctx, &tv, d->v, d->kv_ld, d->kv_bs, d->tk, d->heads, d->batch));
  TS_TRY_RC(head_tmap(ctx, &tdo, d->d_o, d->o_ld, d->o_bs, d->tq, d->heads, d->batch));
  TS_TRY_RC(head_tmap(ctx, &tq64, d->q, d->q_ld, d->q_bs, d->tq, d->heads, d->batch, 64));
  TS_TRY_RC(head_tmap(ctx, &tk64, d->k, d->kv_ld, d->kv_bs, d->tk, d->heads, d->batch, 64));
  TS_TRY_RC(head_tmap(ctx, &tv64, d->v, d->kv_ld, d->kv_bs, d->tk, d->heads, d->batch, 64));
  TS_TRY_RC(head_tmap(ctx, &tdo64, d->d_o, d->o_ld, d->o_bs, d->tq, d->heads, d->batch, 64));
  static bool attr = false;
  if (!attr) {
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_bwd_dq_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem));
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_bwd_dq_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem));
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_bwd_dkv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem));
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_bwd_dkv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem));
    attr = true;
  }
  dim3 gq(cdiv(d->tq, AT_M), d->heads, d->batch), gk(cdiv(d->tk, AT_N), d->heads, d->batch);
  if (d->mask_mode == 0) {
    ts::launch_k(attn_bwd_dq_kernel<0>, gq, kBwdThreads, kDqSmem, st, tq, tk64, tv64, tdo, p, ctx->d_watchdog);
    TS_LAUNCH_OK(ctx);
    ts::launch_k(attn_bwd_dkv_kernel<0>, gk, kBwdThreads, kDkvSmem, st, tq64, tk, tv, tdo64, p, ctx->d_watchdog);
  } else {
    ts::launch_k(attn_bwd_dq_kernel<1>, gq, kBwdThreads, kDqSmem, st, tq, tk64, tv64, tdo, p, ctx->d_watchdog);
    TS_LAUNCH_OK(ctx);
    ts::launch_k(attn_bwd_dkv_kernel<1>, gk, kBwdThreads, kDkvSmem, st, tq64, tk, tv, tdo64, p, ctx->d_watchdog);
  }
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts

extern "C" {
int ts_attn_fwd(ts_ctx* ctx, const ts_attn_desc* d, void* stream) {
  if (!ctx || !d) return TS_EINVAL;
  return ts::attn_fwd(reinterpret_cast<ts::Ctx*>(ctx), d, reinterpret_cast<cudaStream_t>(stream));
}
int ts_attn_bwd(ts_ctx* ctx, const ts_attn_desc* d, void* stream) {
  if (!ctx || !d) return TS_EINVAL;
  return ts::attn_bwd(reinterpret_cast<ts::Ctx*>(ctx), d, reinterpret_cast<cudaStream_t>(stream));
}
}
