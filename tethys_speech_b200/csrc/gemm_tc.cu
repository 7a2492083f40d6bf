// K9 — tcgen05 GEMM engine (bf16 in, fp32 accumulate in TMEM, fused epilogue), persistent and warp-specialised.
//
// grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, then m, batch,
// K-split) so that neighbouring CTAs share the same A rows in L2. Roles (320 threads):
//   warp 0    : TMA producer   (cp.async.bulk.tensor 4D boxes -> 128B-swizzled smem ring of 4 / 6 / 8 stages, mbarrier tx)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16, BN = 64 / 128 / 256)
//   warps 2-9 : epilogue: warp % 4 = TMEM lane quarter (32 rows), (warp - 2) / 4 = which alternate 128-byte column chunks.
//               tcgen05.ld -> registers -> alpha, bias, [pre-activation copy], GELU or GELU', dropout, residual in registers
//               -> 128B-swizzled 32-row staging box in smem -> cp.async.bulk.tensor store (reduce-add for fp32 "C +=").
//               The pre-activation copy is a second TMA store out of the same box; the activation / dropout math runs
//               between issuing it and waiting for it. Outputs TMA cannot describe (unaligned C) take a generic path
//               (32 x 32 transpose through the box, plain coalesced stores).
// (A 16-warp epilogue — two warps per staging box, 32 columns per thread, 96 registers — was built and measured in round 2: correct,
// but 3-4 % SLOWER on every shape (fc1 + GELU + dropout 38.7 -> 40.1 us, conv1 226 -> 236 us): the pair barriers around the shared
// box and the tighter register budget cost more than the extra warps hide. Eight warps it stays.)
// The accumulator is double buffered in TMEM (2 x BN columns): the epilogue of tile i overlaps the main loop of tile
// i+1 (tmem_full / tmem_empty mbarriers).
// Split-K (only for fp32 "C += A*B" outputs, i.e. weight gradients with a long reduce dim and few output tiles):
// each K-slice is its own tile and adds its partial with the TMA reduce (or red.global.add.f32 on the generic path).
// Both operands may be K-major or MN-major (see include/tethys.h); MN-major operands are loaded as
// 64-wide MN chunks so Dense kernels [in,out], activations for wgrad and V for P.V need no transposes.
// Tile width and split factor come from a cost model per shape (gemm_tc(): waves x k-blocks x BN x efficiency + tail).
// Replaces cuBLAS/cuDNN calls behind W:89-92,141,147,167,174,194-205,311-312,545 and V:240-268,
// 316-319,338-348,362,371,383-398 (+ their autodiff transposes).
#include <unordered_map>
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"

namespace ts {

struct EpiParams {
  void* c;
  void* c_pre;
  const void* res;
  const float* bias;
  const void* aux;   // act == 2: GELU input u (same dtype / batch strides as C), row stride ld_aux
  long long ld_aux;
  long long ldc, ldr, c_bs1, c_bs2, r_bs1, r_bs2, bias_bs1;
  float alpha;
  int act, accumulate;
  int m, n, k, nb1;
  int a_m1, a_m2, b_m1, b_m2;  // 0 => that batch dim is broadcast for the operand (stride 0)
  uint32_t drop_thr; float inv_keep; unsigned long long seed;
  const unsigned long long* salt;                     // device-resident dropout salt (Ctx::d_state)
  int mt, nt, nb, splitk, kb_per_split, total_tiles;  // persistent tile schedule
  int use_red;                                        // fp32 output: add with red.global (split-K partials)
  int tma_epi;                                        // C (and c_pre) are TMA-storable: swizzled smem box + bulk tensor store
  unsigned long long* trace;                          // debug (ts_debug_gemm_trace): 16 globaltimer stamps per CTA, else NULL
  double* gn_accum;                                   // GroupNorm statistics in the epilogue (ts_gemm_desc.gn_accum), else NULL
  int raster;                                         // tile walk order (decode_tile)
  int gn_rpb, gn_valid, gn_groups, gn_cpg;            // rows per batch block, data rows of it, groups, channels per group
};
__device__ __forceinline__ void trace_stamp(const EpiParams& p, int slot) {
  if (p.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[blockIdx.x * 16 + slot] = t;
  }
}

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;                 // two warps per TMEM lane quarter, alternating column chunks
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kStageBytesPerWarp = 4096;     // one 32-row x 128-byte swizzled box per epilogue warp

// CTAS = 1: one CTA owns a 128 x BN tile. CTAS = 2: a CTA pair (cluster of two, cta_group::2) owns a 256 x BN tile; each CTA
// stages its own 128 rows of A and its own BN/2 rows of B, the leader's single thread issues 256-row UMMAs that read both
// CTAs' smem and write both CTAs' TMEM, so per FLOP each SM reads half as much B from smem / L2.
template <int BN, int CTAS> struct TcCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / (CTAS == 1 ? 1 : 2)) * BK * 2;   // a CTA of a pair holds half of the tile's B columns
  static constexpr int kStages = (196608 / (kABytes + kBBytes)) > 8 ? 8 : (196608 / (kABytes + kBBytes));
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmem = kStages * kStageBytes + kEpiWarps * kStageBytesPerWarp + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;                     // two accumulator stages (power of two >= 32)
};

struct TileCoord { int m0, n0, b1, b2, kb0, kb1; };
__device__ __forceinline__ TileCoord decode_tile(const EpiParams& p, int t, int BN_, int BM_) {
  TileCoord tc;
  int m, n;
  if (p.raster == 0) { n = t % p.nt; t /= p.nt; m = t % p.mt; t /= p.mt; }   // n fastest: neighbouring CTAs share A rows
  else { m = t % p.mt; t /= p.mt; n = t % p.nt; t /= p.nt; }                 // m fastest: neighbouring CTAs share B columns
  const int b = t % p.nb; t /= p.nb;
  tc.m0 = m * BM_; tc.n0 = n * BN_;
  tc.b1 = b % p.nb1; tc.b2 = b / p.nb1;
  const int nkb = (p.k + BK - 1) / BK;
  tc.kb0 = t * p.kb_per_split;
  tc.kb1 = min(nkb, tc.kb0 + p.kb_per_split);
  return tc;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 16-byte chunk j (0..7) of row r inside a 32-row x 128-byte SWIZZLE_128B box
__device__ __forceinline__ uint32_t swz128(uint32_t box, int r, int j) { return box + r * 128 + ((j ^ (r & 7)) << 4); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Per-element epilogue math on a thread's NV consecutive columns of one output row.
template <int NV>
__device__ __forceinline__ void epi_bias(float (&v)[NV], const EpiParams& p, const float* bias, int col, bool full) {
  if (bias) {   // alpha and the bias in one FMA per element
    const float a = p.alpha;
    if (full && (reinterpret_cast<uintptr_t>(bias + col) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + i));  // warp-uniform address: one broadcast
        v[i] = fmaf(v[i], a, b.x); v[i + 1] = fmaf(v[i + 1], a, b.y); v[i + 2] = fmaf(v[i + 2], a, b.z); v[i + 3] = fmaf(v[i + 3], a, b.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = fmaf(v[i], a, col + i < p.n ? __ldg(bias + col + i) : 0.f);
    }
  } else if (p.alpha != 1.f) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= p.alpha;
  }
}
template <int NV, typename OutT>
__device__ __forceinline__ void epi_act_drop(float (&v)[NV], const EpiParams& p, unsigned long long seed, unsigned long long e0, const OutT* aux_row, bool row_ok,
                                             int col, bool full) {
  if (p.act == 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = gelu_fast_f(v[i]);   // bf16 operands: fast erf (common.cuh)
  } else if (p.act == 2 && row_ok) {
    // backward of a GELU: multiply by GELU'(u) with u read from the kept pre-activation tensor
    if (full && (reinterpret_cast<uintptr_t>(aux_row) & 15) == 0) {
      if (sizeof(OutT) == 4) {
#pragma unroll
        for (int j = 0; j < NV / 4; ++j) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(aux_row) + j);
          v[4 * j] *= gelu_grad_fast_f(f.x); v[4 * j + 1] *= gelu_grad_fast_f(f.y);
          v[4 * j + 2] *= gelu_grad_fast_f(f.z); v[4 * j + 3] *= gelu_grad_fast_f(f.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NV / 8; ++j) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(aux_row) + j);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            v[8 * j + 2 * e] *= gelu_grad_fast_f(f.x);
            v[8 * j + 2 * e + 1] *= gelu_grad_fast_f(f.y);
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (col + i < p.n) v[i] *= gelu_grad_fast_f(to_f<OutT>(aux_row[i]));
    }
  }
  if (p.drop_thr) {
    const DropKey key = flat_drop_key(seed, p.drop_thr);
    if ((e0 & 31) == 0 && NV >= 32) {   // this thread's NV consecutive elements are whole 32-element chunks of the mask stream
#pragma unroll
      for (int c = 0; c < NV / 32; ++c) dropout_apply_chunk(v + 32 * c, key, (uint32_t)(e0 >> 5) + c, p.inv_keep);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] *= dropout_scale(key, e0 + i, p.inv_keep);
    }
  }
}

// GroupNorm statistics of the tile in flight (V:167-176: moments over time x channels of a group, per batch element): the thread
// owns NV consecutive channels of output row `row`; channel sub-chunks of 32 never straddle a group (host check: cpg % 32 == 0).
// A warp's 32 rows normally lie in one batch element: one shuffle reduction, then lane 0 adds the two sums with fp64 atomics
// (~4 atomics per warp and 64-column chunk). A warp that straddles a batch boundary falls back to per-lane atomics.
template <int NV>
__device__ __forceinline__ void gn_epilogue_stats(const float (&v)[NV], const EpiParams& p, int row, int col, int lane) {
  const int bb = row / p.gn_rpb;
  const bool valid = row < p.m && row - bb * p.gn_rpb < p.gn_valid;
  const int b_first = __shfl_sync(0xffffffffu, bb, 0), b_last = __shfl_sync(0xffffffffu, bb, 31);
  const bool uniform = b_first == b_last;
#pragma unroll
  for (int j = 0; j < NV / 32; ++j) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { s += v[32 * j + i]; ss = fmaf(v[32 * j + i], v[32 * j + i], ss); }
    if (!valid) { s = 0.f; ss = 0.f; }
    const int g = (col + 32 * j) / p.gn_cpg;
    if (uniform) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
      if (lane == 0) {
        double* a = p.gn_accum + ((long long)bb * p.gn_groups + g) * 2;
        atomicAdd(a, (double)s); atomicAdd(a + 1, (double)ss);
      }
    } else if (valid) {
      double* a = p.gn_accum + ((long long)bb * p.gn_groups + g) * 2;
      atomicAdd(a, (double)s); atomicAdd(a + 1, (double)ss);
    }
  }
}

template <int BN, int AMAJ, int BMAJ, typename OutT, int CTAS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_p,
               const EpiParams p, int* watchdog) {
  using Cfg = TcCfg<BN, CTAS>;
  static_assert(CTAS == 1 || (CTAS == 2 && BN >= 128) || (CTAS == 4 && BN == 256), "a CTA pair needs >= 64 B rows per CTA; the 4-CTA form 256 columns");
  // CTAS = 2: a CTA pair; the even rank (leader) issues the MMAs; tiles are walked per cluster.
  // CTAS = 4: a cluster of TWO pairs on vertically adjacent 256-row tiles of the same column block: the B operand is the same for both
  // pairs, so every CTA fetches only a quarter of it (64 rows) and TMA-multicasts it to its counterpart in the other pair — per SM and
  // k-block 16 KB of A + 8 KB of B come out of L2 instead of 16 + 16 (operand delivery is what bounds these main loops, gemm_tc()).
  constexpr bool PAIR = CTAS >= 2, MC = CTAS == 4;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;   // 0 .. CTAS-1
  const uint32_t rp = rank & 1u;                               // rank inside the pair
  const uint32_t pr = rank >> 1;                               // pair inside the cluster
  const uint32_t lead = rank & ~1u;                            // cluster rank of this pair's leader
  const int tile_first = PAIR ? (int)(blockIdx.x / CTAS) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x / CTAS) : (int)gridDim.x;
  constexpr int BMT = BM * CTAS;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_buf = smem + S * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_buf + kEpiWarps * kStageBytesPerWarp);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tmem_full_bar = empty_bar + S;       // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_stamp(p, 0);   // kernel entry

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], MC ? 2 : 1);   // 4-CTA form: a slot is free when BOTH pairs' MMAs have retired (each multicasts into it)
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full_bar[s], 1);
      ptx::mbar_init(&tmem_empty_bar[s], kEpiWarps * (PAIR ? 2 : 1));   // pair: the leader's barrier collects both CTAs' epilogue warps
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { ptx::tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols); ptx::tmem_relinquish_2cta(); }
    else { ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  // The producer warp of a (leader) CTA needs nothing but the barriers its own lane 0 has just initialised: it only ARRIVES at the
  // setup barrier and starts the first TMA loads while warp 1 is still allocating TMEM (the head of a launch is not overlapped
  // by anything else). Everyone else waits; a peer CTA's producer signals the LEADER's barriers and so must wait for the cluster.
  // (4-CTA form: every producer multicasts into, and signals barriers of, other CTAs: all of them wait for the cluster.)
  const bool early = (warp == 0 && rank == 0 && !MC);
  if (PAIR) {
    ptx::cluster_arrive();
    if (!early) ptx::cluster_wait();
  } else {
    if (early) ptx::named_bar_arrive(1, kThreads);
    else ptx::named_bar_sync(1, kThreads);
  }
  ptx::tc_fence_after();
  const uint32_t tmem_base = early ? 0u : *tmem_slot;   // warp 0 never touches TMEM
  ts::pdl_enter();   // prologue above overlaps the previous grid's tail (PDL, common.cuh)
  if (threadIdx.x == 32) trace_stamp(p, 1);   // setup done (barriers, TMEM, CTA / cluster sync)

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      bool ok = true;
      for (int t = tile_first; t < p.total_tiles && ok; t += tile_step) {
        const TileCoord tc = decode_tile(p, t, BN, BMT);
        // this CTA's rows of A, and the rows of B it fetches: its half of the tile's columns (pair), or the quarter it multicasts (4-CTA)
        const int m0 = tc.m0 + (int)rank * BM, n0 = tc.n0 + (int)rp * (BN / 2) + (MC ? (int)pr * (BN / 4) : 0);
        for (int kb = tc.kb0; kb < tc.kb1; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          if (!ptx::mbar_wait(&empty_bar[s], ph ^ 1, watchdog, 1)) { ok = false; break; }
          // pair: all bytes of the stage (both CTAs' loads) are counted on the LEADER's barrier
          const uint32_t fb = PAIR ? ptx::mapa_u32(ptx::smem_u32(&full_bar[s]), lead) : 0u;
          if (rp == 0) ptx::mbar_expect_tx(&full_bar[s], Cfg::kStageBytes * (PAIR ? 2 : 1));   // what lands in this pair's two CTAs
          auto load = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
            if (PAIR) ptx::tma_load_4d_2cta(dst, tm, fb, c0, c1, c2, c3);
            else ptx::tma_load_4d(dst, tm, &full_bar[s], c0, c1, c2, c3);
          };
          // 4-CTA form: the same smem offset in this CTA and in its counterpart (same rank-in-pair) of the other pair; the
          // transaction bytes are counted on each destination pair's leader barrier (cta_group::2 barrier addressing)
          const uint16_t mc_mask = (uint16_t)((1u << rp) | (1u << (2 + rp)));
          auto load_mc = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
            ptx::tma_load_4d_2cta_mc(dst, tm, fb, c0, c1, c2, c3, mc_mask);
          };
          const uint32_t sa = ptx::smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          if (AMAJ == 0) {
            load(sa, &tma_a, kb * BK, m0, tc.b1 * p.a_m1, tc.b2 * p.a_m2);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              load(sa + c * 8192, &tma_a, m0 + c * 64, kb * BK, tc.b1 * p.a_m1, tc.b2 * p.a_m2);
          }
          if (MC) {   // 64 rows (K-major box) / one 64-wide chunk (MN-major) of B: 8 KB at offset pr * 8 KB of the B half
            if (BMAJ == 0) load_mc(sb + pr * 8192, &tma_b, kb * BK, n0, tc.b1 * p.b_m1, tc.b2 * p.b_m2);
            else load_mc(sb + pr * 8192, &tma_b, n0, kb * BK, tc.b1 * p.b_m1, tc.b2 * p.b_m2);
          } else if (BMAJ == 0) {
            load(sb, &tma_b, kb * BK, n0, tc.b1 * p.b_m1, tc.b2 * p.b_m2);
          } else {
#pragma unroll
            for (int c = 0; c < BN / (PAIR ? 2 : 1) / 64; ++c)
              load(sb + c * 8192, &tma_b, n0 + c * 64, kb * BK, tc.b1 * p.b_m1, tc.b2 * p.b_m2);
          }
        }
      }
    }
    if (CTAS == 2 && rank == 0) { __syncwarp(); ptx::cluster_wait(); }   // second half of the setup barrier (completed long ago); early producer only
    __syncwarp();
    ts::pdl_tail();   // every operand load of this CTA is in flight: let the next grid's CTAs take the SMs as they free up
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && rp == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 2 * BM : BM, BN, AMAJ, BMAJ);
      const uint16_t pair_mask = (uint16_t)(3u << lead);
      auto commit = [&](uint64_t* bar) {          // accumulator ready: both CTAs of this pair
        if (PAIR) ptx::umma_commit_2cta(bar, pair_mask);
        else ptx::umma_commit(bar);
      };
      auto commit_slot = [&](uint64_t* bar) {     // operand slot free: every CTA that writes into it (4-CTA form: the whole cluster)
        if (PAIR) ptx::umma_commit_2cta(bar, MC ? (uint16_t)0xF : pair_mask);
        else ptx::umma_commit(bar);
      };
      uint32_t it = 0, tl = 0;
      bool ok = true;
      for (int t = tile_first; t < p.total_tiles && ok; t += tile_step, ++tl) {
        const TileCoord tc = decode_tile(p, t, BN, BMT);
        const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
        if (!ptx::mbar_wait(&tmem_empty_bar[as], aph ^ 1, watchdog, 4)) { ok = false; break; }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = tc.kb0; kb < tc.kb1; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          if (!ptx::mbar_wait(&full_bar[s], ph, watchdog, 2)) { ok = false; break; }
          if (it == 0) trace_stamp(p, 2);    // first operand stage landed
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adesc = (AMAJ == 0) ? ptx::make_smem_desc(sa + kk * 32, 16, 1024)
                                               : ptx::make_smem_desc(sa + kk * 2048, 8192, 1024);
            const uint64_t bdesc = (BMAJ == 0) ? ptx::make_smem_desc(sb + kk * 32, 16, 1024)
                                               : ptx::make_smem_desc(sb + kk * 2048, 8192, 1024);
            if (PAIR) ptx::umma_f16_2cta(d_tmem, adesc, bdesc, idesc, (kb > tc.kb0 || kk > 0) ? 1u : 0u);
            else ptx::umma_f16(d_tmem, adesc, bdesc, idesc, (kb > tc.kb0 || kk > 0) ? 1u : 0u);
          }
          commit_slot(&empty_bar[s]);  // frees the smem slot (in both CTAs of a pair / all four of the cluster) when these MMAs retire
        }
        if (ok) commit(&tmem_full_bar[as]);  // accumulator of this tile complete
        if (tl == 0) trace_stamp(p, 3);      // first tile's MMAs issued
      }
      trace_stamp(p, 4);                     // last MMA issued
    }
  } else {
    // ===== epilogue: warps 2..9; warp % 4 = TMEM lane quarter, (warp - 2) / 4 = which alternate column chunks =====
    constexpr bool kF32 = sizeof(OutT) == 4;
    constexpr int CW = kF32 ? 32 : 64;  // columns per chunk = one 128-byte staging row
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    uint8_t* wbuf = stage_buf + ew * kStageBytesPerWarp;
    const unsigned long long seed = p.drop_thr ? salted_seed(p.seed, p.salt) : p.seed;
    const uint32_t wbuf_s = ptx::smem_u32(wbuf);
    uint32_t tl = 0;
    for (int t = tile_first; t < p.total_tiles; t += tile_step, ++tl) {
      TileCoord tc = decode_tile(p, t, BN, BMT);
      tc.m0 += (int)rank * BM;   // this CTA's 128 rows of the pair tile (its own TMEM lanes)
      const uint32_t as = tl & 1, aph = (tl >> 1) & 1;
      if (!ptx::mbar_wait(&tmem_full_bar[as], aph, watchdog, 3)) break;
      if (ew == 0 && lane == 0) trace_stamp(p, t + tile_step >= p.total_tiles ? 6 : 5);   // accumulator ready: a tile / the last tile
      ptx::tc_fence_after();
      const long long boff = (long long)tc.b1 * p.c_bs1 + (long long)tc.b2 * p.c_bs2;
      const float* bias = p.bias ? p.bias + (long long)tc.b1 * p.bias_bs1 : nullptr;
      const int row0 = tc.m0 + q * 32;
      const int row = row0 + lane;
      const uint32_t tm = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
      bool released = false;
      auto release = [&]() {
        if (!released) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tmem_empty_bar[as]), lead));
            else ptx::mbar_arrive(&tmem_empty_bar[as]);
          }
          released = true;
        }
      };
      if (row0 >= p.m) { release(); continue; }  // this warp's 32 rows are all padding
      if (p.tma_epi) {
        // ---- registers -> swizzled smem box -> TMA store / reduce-add (coalesced, clipped at the m/n edges) ----
#pragma unroll 1
        for (int c0 = half * CW; c0 < BN; c0 += 2 * CW) {
          const int col = tc.n0 + c0;
          if (col >= p.n) break;
          float v[CW];
          {
            uint32_t r[32];
            ptx::tmem_ld_32x32(tm + (uint32_t)c0, r);
            if constexpr (CW == 64) {
              uint32_t r2[32];
              ptx::tmem_ld_32x32(tm + (uint32_t)c0 + 32, r2);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[32 + i] = __uint_as_float(r2[i]);
            } else {
              ptx::tmem_ld_wait();
            }
            if (ew == 0 && lane == 0 && c0 == 0 && t + tile_step >= p.total_tiles) trace_stamp(p, 8);   // first chunk of the last tile in registers
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          }
          if (c0 + 2 * CW >= BN || tc.n0 + c0 + 2 * CW >= p.n) release();  // last TMEM read of this warp for this tile
          const bool full = col + CW <= p.n;
          epi_bias<CW>(v, p, bias, col, full);
          if (p.gn_accum) gn_epilogue_stats<CW>(v, p, row, col, lane);
          // the previous TMA store of this warp must have finished reading the staging box
          if (lane == 0) ptx::bulk_wait_read<0>();
          __syncwarp();
          if (p.c_pre) {
            if (kF32) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts128(swz128(wbuf_s, lane, j), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts128(swz128(wbuf_s, lane, j), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_4d(&tma_p, wbuf_s, col, row0, tc.b1, tc.b2);
              ptx::bulk_commit();
            }
          }
          // activation / dropout work on registers only: it overlaps the TMA engine reading the pre-activation box
          epi_act_drop<CW, OutT>(v, p, seed, (unsigned long long)(boff + (long long)row * p.ldc + col),
                                 reinterpret_cast<const OutT*>(p.aux) + boff + (long long)row * p.ld_aux + col, row < p.m, col, full);
          if (p.res && row < p.m) {
            const OutT* rrow = reinterpret_cast<const OutT*>(p.res) + (long long)tc.b1 * p.r_bs1 + (long long)tc.b2 * p.r_bs2 +
                               (long long)row * p.ldr + col;
            if (full && (reinterpret_cast<uintptr_t>(rrow) & 15) == 0) {
              if (kF32) {
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                  const float4 f = __ldg(reinterpret_cast<const float4*>(rrow) + j);
                  v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < CW / 8; ++j) {
                  const uint4 u = __ldg(reinterpret_cast<const uint4*>(rrow) + j);
                  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h[e]);
                    v[8 * j + 2 * e] += f.x; v[8 * j + 2 * e + 1] += f.y;
                  }
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < CW; ++i)
                if (col + i < p.n) v[i] += to_f<OutT>(rrow[i]);
            }
          }
          if (p.c_pre) {   // the pre-activation store must be done with the staging box before it is overwritten
            if (lane == 0) ptx::bulk_wait_read<0>();
            __syncwarp();
          }
          if (kF32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(swz128(wbuf_s, lane, j), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(swz128(wbuf_s, lane, j), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                     pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (ew == 0 && lane == 0 && t + tile_step >= p.total_tiles) trace_stamp(p, c0 == 0 ? 9 : 10);   // chunk staged (first / later)
          if (lane == 0) {
            if (p.accumulate) ptx::tma_reduce_add_4d(&tma_c, wbuf_s, col, row0, tc.b1, tc.b2);
            else ptx::tma_store_4d(&tma_c, wbuf_s, col, row0, tc.b1, tc.b2);
            ptx::bulk_commit();
          }
        }
        release();
      } else {
        // ---- generic path (unaligned C): 32x32 transpose through the warp's smem box, coalesced plain stores ----
        float* sbuf = reinterpret_cast<float*>(wbuf);
        OutT* cbase = reinterpret_cast<OutT*>(p.c) + boff;
        OutT* pbase = p.c_pre ? reinterpret_cast<OutT*>(p.c_pre) + boff : nullptr;
        const OutT* rbase = p.res ? reinterpret_cast<const OutT*>(p.res) + (long long)tc.b1 * p.r_bs1 + (long long)tc.b2 * p.r_bs2 : nullptr;
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          if (tc.n0 + c0 >= p.n) break;
          uint32_t r[32];
          ptx::tmem_ld_32x32(tm + (uint32_t)c0, r);
          ptx::tmem_ld_wait();
          if (c0 + 64 >= BN || tc.n0 + c0 + 64 >= p.n) release();
#pragma unroll
          for (int j = 0; j < 32; ++j) sbuf[lane * 32 + (j ^ lane)] = __uint_as_float(r[j]);
          __syncwarp();
          const int col = tc.n0 + c0 + lane;   // lane = column
          if (col < p.n) {
            const float bv = bias ? __ldg(bias + col) : 0.f;
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
              const int grow = row0 + rr;
              if (grow >= p.m) break;
              float v = sbuf[rr * 32 + (lane ^ rr)] * p.alpha + bv;
              const long long off = (long long)grow * p.ldc + col;
              if (pbase) pbase[off] = from_f<OutT>(v);
              if (p.act == 1) v = gelu_fast_f(v);
              else if (p.act == 2) v *= gelu_grad_fast_f(to_f<OutT>(reinterpret_cast<const OutT*>(p.aux)[boff + (long long)grow * p.ld_aux + col]));
              if (p.drop_thr) v *= dropout_scale(seed, (unsigned long long)(boff + off), p.drop_thr, p.inv_keep);
              if (rbase) v += to_f<OutT>(rbase[(long long)grow * p.ldr + col]);
              if (kF32) {
                float* dst = reinterpret_cast<float*>(cbase) + off;
                if (p.use_red) atomicAdd(dst, v);
                else if (p.accumulate) *dst += v;
                else *dst = v;
              } else {
                if (p.accumulate) v += to_f<OutT>(cbase[off]);
                cbase[off] = from_f<OutT>(v);
              }
            }
          }
          __syncwarp();
        }
        release();
      }
    }
    // the bulk stores of this warp must have READ their staging box before the CTA exits (smem is released at exit); their global
    // writes complete by the end of the grid like any other store
    if (lane == 0) ptx::bulk_wait_read<0>();
    if (ew == 0 && lane == 0) trace_stamp(p, 7);   // this warp's stores have drained
  }
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync_all();   // the leader's MMAs read the peer's smem / write its TMEM until the last commit
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) ptx::tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
    else ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static int resolve_encode(Ctx* ctx) {
  if (ctx->encode_tiled) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
    return set_err(ctx, TS_ECUDA, "cuTensorMapEncodeTiled entry point unavailable (%s)",
                   cudaGetErrorString(e));
  ctx->encode_tiled = fn;
  return 0;
}

struct TmapKey {
  const void* base;
  uint64_t d[4];
  uint64_t s[3];
  uint32_t box[2];
  uint64_t dtype;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return h;
  }
};
typedef std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> TmapCache;

void tmap_cache_free(Ctx* ctx) {
  if (ctx->tmap_cache) delete reinterpret_cast<TmapCache*>(ctx->tmap_cache);
  ctx->tmap_cache = nullptr;
}

// bf16 tensor map: dims d[0..3] (d[0] innermost, contiguous), strides in BYTES for dims 1..3.
int get_tmap(Ctx* ctx, CUtensorMap* out, const void* base, const uint64_t d[4], const uint64_t sbytes[3],
             uint32_t box0, uint32_t box1, bool f32 = false) {
  if (resolve_encode(ctx)) return TS_ECUDA;
  if (!ctx->tmap_cache) ctx->tmap_cache = new TmapCache();
  TmapCache& cache = *reinterpret_cast<TmapCache*>(ctx->tmap_cache);
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  for (int i = 0; i < 4; ++i) key.d[i] = d[i];
  for (int i = 0; i < 3; ++i) key.s[i] = sbytes[i];
  key.box[0] = box0;
  key.box[1] = box1;
  key.dtype = f32 ? 1 : 0;
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return 0; }
  cuuint64_t gd[4] = {d[0], d[1], d[2], d[3]};
  cuuint64_t gs[3] = {sbytes[0], sbytes[1], sbytes[2]};
  cuuint32_t bx[4] = {box0, box1, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(ctx, TS_ECUDA,
                   "cuTensorMapEncodeTiled failed (%d): base=%p dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu] "
                   "box=[%u,%u]",
                   (int)r, base, (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)d[2],
                   (unsigned long long)d[3], (unsigned long long)sbytes[0], (unsigned long long)sbytes[1],
                   (unsigned long long)sbytes[2], box0, box1);
  if (cache.size() > 65536) cache.clear();
  cache[key] = *out;
  return 0;
}

// Can this GEMM run on the TMA/tcgen05 engine?  (16-byte aligned bases and strides.)
bool gemm_tc_supported(const ts_gemm_desc* d) {
  if (d->in_dtype != TS_BF16) return false;
  if (d->out_dtype != TS_BF16 && d->out_dtype != TS_F32) return false;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  auto s16 = [](long long elems) { return (elems * 2) % 16 == 0 && elems > 0; };
  if (!al(d->a) || !al(d->b)) return false;
  if (!s16(d->lda) || !s16(d->ldb)) return false;
  auto bs_ok = [&](long long e) { return e == 0 || s16(e); };
  if (d->batch1 > 1 && (!bs_ok(d->a_bs1) || !bs_ok(d->b_bs1))) return false;
  if (d->batch2 > 1 && (!bs_ok(d->a_bs2) || !bs_ok(d->b_bs2))) return false;
  if (d->m <= 0 || d->n <= 0 || d->k <= 0) return false;
  return true;
}

template <int BN, int AMAJ, int BMAJ, typename OutT, int CTAS>
static int launch_tc(Ctx* ctx, const ts_gemm_desc* d, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                     const CUtensorMap& tp, const EpiParams& ep, cudaStream_t st) {
  using Cfg = TcCfg<BN, CTAS>;
  auto kern = gemm_tc_kernel<BN, AMAJ, BMAJ, OutT, CTAS>;
  static bool attr_set = false;
  if (!attr_set) {
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_set = true;
  }
  if (CTAS == 1) {
    const int grid = ep.total_tiles < ctx->num_sms ? ep.total_tiles : ctx->num_sms;
    ts::launch_k(kern, grid, kThreads, Cfg::kSmem, st, ta, tb, tc, tp, ep, ctx->d_watchdog);
  } else {
    // one cluster of two CTAs (one TPC) per 256-row tile stream, or of four (two TPCs of one GPC) per 512-row stream
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads, 1, 1); cfg.dynamicSmemBytes = Cfg::kSmem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    int units = ctx->num_sms / CTAS;
    if (CTAS == 4) {   // clusters of four do not tile every GPC: size the persistent grid to what is co-resident
      static int max_clusters = 0;
      if (!max_clusters) {
        cfg.gridDim = dim3(CTAS * units, 1, 1); cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = units * 3 / 4; }
        max_clusters = n;
      }
      if (units > max_clusters) units = max_clusters;
    }
    const int grid = CTAS * (ep.total_tiles < units ? ep.total_tiles : units);
    cfg.gridDim = dim3(grid, 1, 1);
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    int* wd = ctx->d_watchdog;
    TS_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, tp, ep, wd));
  }
  TS_LAUNCH_OK(ctx);
  return 0;
}

template <int BN, typename OutT, int CTAS>
static int dispatch_major(Ctx* ctx, const ts_gemm_desc* d, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                          const CUtensorMap& tp, const EpiParams& ep, cudaStream_t st) {
  if (d->a_major == 0 && d->b_major == 0) return launch_tc<BN, 0, 0, OutT, CTAS>(ctx, d, ta, tb, tc, tp, ep, st);
  if (d->a_major == 0 && d->b_major == 1) return launch_tc<BN, 0, 1, OutT, CTAS>(ctx, d, ta, tb, tc, tp, ep, st);
  if (d->a_major == 1 && d->b_major == 0) return launch_tc<BN, 1, 0, OutT, CTAS>(ctx, d, ta, tb, tc, tp, ep, st);
  return launch_tc<BN, 1, 1, OutT, CTAS>(ctx, d, ta, tb, tc, tp, ep, st);
}

int gemm_tc(Ctx* ctx, const ts_gemm_desc* d, cudaStream_t st) {
  TS_REQUIRE(ctx, gemm_tc_supported(d), TS_EUNSUPPORTED,
             "gemm_tc: operands must be bf16 with 16-byte aligned bases/strides");
  TS_REQUIRE(ctx, !(d->accumulate && (d->act || d->residual)), TS_EINVAL, "gemm: accumulate excludes act/residual");
  TS_REQUIRE(ctx, d->act != 2 || d->act_aux, TS_EINVAL, "gemm: act 2 needs act_aux");
  if (d->gn_accum)
    TS_REQUIRE(ctx, !d->act && !d->residual && !d->accumulate && !d->c_preact && d->drop <= 0.f && d->batch1 <= 1 && d->batch2 <= 1 &&
                   d->gn_groups > 0 && d->n % d->gn_groups == 0 && (d->n / d->gn_groups) % 32 == 0 && d->gn_rows_per_batch > 0 &&
                   d->gn_valid_rows > 0 && d->gn_valid_rows <= d->gn_rows_per_batch,
               TS_EINVAL, "gemm: gn_accum needs a plain epilogue, no batching and n / gn_groups a multiple of 32 (n=%lld groups=%d)",
               (long long)d->n, (int)d->gn_groups);
  const int nb1 = d->batch1 > 0 ? d->batch1 : 1, nb2 = d->batch2 > 0 ? d->batch2 : 1;
  // Tile-N and split-K choice from a small cost model (units: 64-deep k-blocks of one 128 x 1 column strip):
  //   cost = waves * k_blocks_per_work_item * bn * eff(bn) + tail_epilogue * bn
  // Wider tiles read less smem per FLOP (eff), more/smaller work items fill the 148 SMs and shorten the un-overlapped
  // last epilogue. Split-K only where partial sums may be added: fp32 "C += A*B" with a plain epilogue (wgrad).
  // A CTA pair (256-row tiles, cta_group::2) halves the B bytes every SM pulls through smem per FLOP (the 1-CTA 128 x 256 tile
  // needs 96 B/clk of operand reads + 96 B/clk of TMA writes against a 128 B/clk smem port); it has half as many work items.
  const int nkb = cdiv(d->k, BK), sms = ctx->num_sms;
  const bool split_ok = d->out_dtype == TS_F32 && d->accumulate && !d->bias && !d->c_preact && d->drop <= 0.f && nkb >= 16;
  static const int force_ctas = getenv("TETHYS_GEMM_CTAS") ? atoi(getenv("TETHYS_GEMM_CTAS")) : 0;
  const int want_ctas = d->force_engine == 3 ? 2 : d->force_engine == 4 ? 4 : force_ctas;
  int bn = 64, splitk = 1, ctas = 1;
  {
    // Cost model in SM cycles, fitted to tools/selftest_gemm timings (profiles/r02e_gemm_tile_sweep.log). What bounds a main loop
    // here is not the tensor pipe but operand delivery: the L2 -> SM path gives an SM ~43 B/clk (B300_MICROARCH: LTS cap ~6300 B/clk
    // chip-wide), a 128 x BN x 64 k-block needs (128 + BN / ctas) * 128 B of operands against 2 * BN MMA cycles, so
    //   k-block cycles = max(2 BN, (16384 + BN / ctas * 128) / 43): 577 (BN 64), 769 (128), 1154 (256), 769 (pair 256, 2x the rows)
    // i.e. per output column 9.0 / 6.0 / 4.5 / 3.0: the widest tile that still fills the SMs wins. Around it: ~1500 cycles per extra
    // wave (tile switch: accumulator hand-over, ring refill), a last epilogue nothing overlaps (~24 cycles per tile column, 1.5x for
    // fp32), ~9000 cycles of cluster set-up and slower ramp for CTA pairs. Split-K only where partial sums may be added (fp32 "C +=", plain epilogue).
    double best = 1e30;
    const int cand[3] = {256, 128, 64};
    static const int force_bn = getenv("TETHYS_GEMM_BN") ? atoi(getenv("TETHYS_GEMM_BN")) : 0;
    for (int cs = 1; cs <= 4; cs *= 2) {
      if (want_ctas && cs != want_ctas) continue;
      if (cs == 4 && want_ctas != 4) continue;          // 4-CTA multicast form: opt-in (TETHYS_GEMM_CTAS=4 / force_engine 4)
      if (cs >= 2 && d->m <= BM * (cs / 2)) continue;
      const int units = sms / cs;                     // CTAs or CTA pairs
      const int mtc = cdiv(d->m, BM * cs);
      for (int i = 0; i < (cs == 4 ? 1 : cs == 2 ? (want_ctas ? 2 : 1) : 3); ++i) {   // pair tiles: 256 wide (128 only when forced: measured slower)
        if (cand[i] > 64 && d->n <= cand[i] / 2) continue;
        if (force_bn && cand[i] != force_bn) continue;
        const long long tiles = (long long)mtc * cdiv(d->n, cand[i]) * nb1 * nb2;
        int sk = 1;
        if (split_ok && tiles < units) {
          sk = (int)(units / tiles);
          const int cap = nkb / 8;
          if (sk > cap) sk = cap;
          if (sk < 1) sk = 1;
        }
        const int kb = cdiv(nkb, sk);
        sk = cdiv(nkb, kb);
        const double waves = (double)((tiles * sk + units - 1) / units);
        const double kbc = fmax(2.0 * cand[i], (16384.0 + (double)cand[i] / cs * 128.0) / 42.6);   // cs 4: a quarter of B per SM
        const double tail = 24.0 * cand[i] * (d->out_dtype == TS_F32 ? 1.5 : 1.0);
        const double score = waves * kb * kbc + (waves - 1.0) * 1500.0 + tail + (cs >= 2 ? 9000.0 : 0.0);
        if (score < best) { best = score; bn = cand[i]; splitk = sk; ctas = cs; }
      }
    }
    TS_REQUIRE(ctx, best < 1e30, TS_EUNSUPPORTED, "gemm_tc: no tile configuration for m=%lld n=%lld (TETHYS_GEMM_CTAS=%d)",
               (long long)d->m, (long long)d->n, want_ctas);
  }
  const int mt = cdiv(d->m, BM * ctas);
  const int nt = cdiv(d->n, bn);
  const long long tiles_ll = (long long)mt * nt * nb1 * nb2;
  TS_REQUIRE(ctx, tiles_ll * splitk < (1ll << 30), TS_ESHAPE, "gemm_tc: too many tiles");
  const int kbps = cdiv(nkb, splitk);

  CUtensorMap ta, tb;
  {
    uint64_t dims[4], str[3];
    const bool bc1 = nb1 <= 1 || d->a_bs1 == 0, bc2 = nb2 <= 1 || d->a_bs2 == 0;
    const uint64_t bs1 = (uint64_t)(bc1 ? d->lda : d->a_bs1) * 2, bs2 = (uint64_t)(bc2 ? d->lda : d->a_bs2) * 2;
    if (d->a_major == 0) { dims[0] = d->k; dims[1] = d->m; } else { dims[0] = d->m; dims[1] = d->k; }
    dims[2] = bc1 ? 1 : nb1; dims[3] = bc2 ? 1 : nb2;
    str[0] = (uint64_t)d->lda * 2; str[1] = bs1; str[2] = bs2;
    int r = get_tmap(ctx, &ta, d->a, dims, str, 64, d->a_major == 0 ? BM : BK);
    if (r) return r;
  }
  {
    uint64_t dims[4], str[3];
    const bool bc1 = nb1 <= 1 || d->b_bs1 == 0, bc2 = nb2 <= 1 || d->b_bs2 == 0;
    const uint64_t bs1 = (uint64_t)(bc1 ? d->ldb : d->b_bs1) * 2, bs2 = (uint64_t)(bc2 ? d->ldb : d->b_bs2) * 2;
    if (d->b_major == 0) { dims[0] = d->k; dims[1] = d->n; } else { dims[0] = d->n; dims[1] = d->k; }
    dims[2] = bc1 ? 1 : nb1; dims[3] = bc2 ? 1 : nb2;
    str[0] = (uint64_t)d->ldb * 2; str[1] = bs1; str[2] = bs2;
    int r = get_tmap(ctx, &tb, d->b, dims, str, 64, d->b_major == 0 ? (uint32_t)(bn / ctas) : BK);
    if (r) return r;
  }
  EpiParams ep;
  ep.c = d->c; ep.c_pre = d->c_preact; ep.res = d->residual; ep.bias = d->bias; ep.aux = d->act_aux; ep.ld_aux = d->ld_aux;
  ep.ldc = d->ldc; ep.ldr = d->ldr; ep.c_bs1 = d->c_bs1; ep.c_bs2 = d->c_bs2; ep.r_bs1 = d->r_bs1; ep.r_bs2 = d->r_bs2; ep.bias_bs1 = d->bias_bs1;
  ep.alpha = d->alpha; ep.act = d->act; ep.accumulate = d->accumulate;
  ep.m = d->m; ep.n = d->n; ep.k = d->k; ep.nb1 = nb1;
  ep.mt = mt; ep.nt = nt; ep.nb = nb1 * nb2; ep.splitk = splitk; ep.kb_per_split = kbps;
  ep.total_tiles = (int)(tiles_ll * splitk); ep.use_red = splitk > 1 ? 1 : 0;
  ep.a_m1 = (nb1 > 1 && d->a_bs1 != 0) ? 1 : 0; ep.a_m2 = (nb2 > 1 && d->a_bs2 != 0) ? 1 : 0;
  ep.b_m1 = (nb1 > 1 && d->b_bs1 != 0) ? 1 : 0; ep.b_m2 = (nb2 > 1 && d->b_bs2 != 0) ? 1 : 0;
  ep.drop_thr = 0; ep.inv_keep = 1.f; ep.seed = d->seed; ep.salt = ctx->d_state;
  ep.trace = reinterpret_cast<unsigned long long*>(ctx->gemm_trace);
  static const int force_raster = getenv("TETHYS_GEMM_RASTER") ? atoi(getenv("TETHYS_GEMM_RASTER")) : -1;
  ep.raster = force_raster >= 0 ? force_raster : 0;
  ep.gn_accum = d->gn_accum; ep.gn_rpb = d->gn_rows_per_batch; ep.gn_valid = d->gn_valid_rows; ep.gn_groups = d->gn_groups;
  ep.gn_cpg = d->gn_accum ? d->n / d->gn_groups : 1;
  if (d->drop > 0.f) {
    double t = (double)d->drop * 4294967296.0;
    ep.drop_thr = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
    ep.inv_keep = 1.f / (1.f - d->drop);
  }
  // TMA-store epilogue when C (and the pre-activation copy) can be described by a tensor map
  CUtensorMap tc, tp;
  memset(&tc, 0, sizeof(tc)); memset(&tp, 0, sizeof(tp));
  {
    const bool f32 = d->out_dtype == TS_F32;
    const uint64_t esz = f32 ? 4 : 2;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    auto s16b = [&](long long elems) { return elems > 0 && ((uint64_t)elems * esz) % 16 == 0; };
    bool ok = al16(d->c) && s16b(d->ldc) && (!d->c_preact || al16(d->c_preact));
    if (nb1 > 1) ok = ok && s16b(d->c_bs1);
    if (nb2 > 1) ok = ok && s16b(d->c_bs2);
    if (!f32 && d->accumulate) ok = false;     // bf16 read-modify-write stays on the generic path
    if (ok) {
      uint64_t dims[4] = {(uint64_t)d->n, (uint64_t)d->m, (uint64_t)nb1, (uint64_t)nb2};
      uint64_t str[3] = {(uint64_t)d->ldc * esz, (uint64_t)(nb1 > 1 ? d->c_bs1 : d->ldc) * esz, (uint64_t)(nb2 > 1 ? d->c_bs2 : d->ldc) * esz};
      int r = get_tmap(ctx, &tc, d->c, dims, str, f32 ? 32 : 64, 32, f32);
      if (r) return r;
      if (d->c_preact) { r = get_tmap(ctx, &tp, d->c_preact, dims, str, f32 ? 32 : 64, 32, f32); if (r) return r; }
    }
    ep.tma_epi = ok ? 1 : 0;
    TS_REQUIRE(ctx, ok || !d->gn_accum, TS_EUNSUPPORTED, "gemm: gn_accum needs a TMA-storable C (16-byte aligned base and row stride)");
    if (ok && splitk > 1) ep.accumulate = 1;   // split-K partials are added by the TMA reduce
  }
  ts_gemm_desc dd = *d;
  dd.batch1 = nb1; dd.batch2 = nb2;
  if (d->out_dtype == TS_BF16) {
    if (ctas == 4) return dispatch_major<256, bf16, 4>(ctx, &dd, ta, tb, tc, tp, ep, st);
    if (ctas == 2) {
      if (bn == 128) return dispatch_major<128, bf16, 2>(ctx, &dd, ta, tb, tc, tp, ep, st);
      return dispatch_major<256, bf16, 2>(ctx, &dd, ta, tb, tc, tp, ep, st);
    }
    if (bn == 64) return dispatch_major<64, bf16, 1>(ctx, &dd, ta, tb, tc, tp, ep, st);
    if (bn == 128) return dispatch_major<128, bf16, 1>(ctx, &dd, ta, tb, tc, tp, ep, st);
    return dispatch_major<256, bf16, 1>(ctx, &dd, ta, tb, tc, tp, ep, st);
  } else {
    if (ctas == 4) return dispatch_major<256, float, 4>(ctx, &dd, ta, tb, tc, tp, ep, st);
    if (ctas == 2) {
      if (bn == 128) return dispatch_major<128, float, 2>(ctx, &dd, ta, tb, tc, tp, ep, st);
      return dispatch_major<256, float, 2>(ctx, &dd, ta, tb, tc, tp, ep, st);
    }
    if (bn == 64) return dispatch_major<64, float, 1>(ctx, &dd, ta, tb, tc, tp, ep, st);
    if (bn == 128) return dispatch_major<128, float, 1>(ctx, &dd, ta, tb, tc, tp, ep, st);
    return dispatch_major<256, float, 1>(ctx, &dd, ta, tb, tc, tp, ep, st);
  }
}

}  // namespace ts
