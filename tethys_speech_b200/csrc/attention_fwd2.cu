// K10 (forward, second generation) — persistent, two query tiles in flight per CTA, one softmax thread per score row.
//
// Why: the first-generation forward (attention_tc.cu, one 128-row query tile per CTA, four threads per row) ran at 0.13 of the
// bf16 peak: 576 short-lived CTAs paid ~6k cycles of prologue each, and every score tile crossed a smem row-max exchange and a
// named barrier. Here one CTA per SM walks work items (batch, head, PAIR of 128-row query tiles); per key/value tile the single
// MMA thread issues S0 = Q0 K^T, S1 = Q1 K^T, O0 += P0 V, O1 += P1 V, and two softmax warp-groups (4 warps = 128 rows each)
// alternate: while group 0 exponentiates S0 the tensor pipe computes S1 / P1 V, and vice versa. A softmax thread owns a whole
// 128-column score row in registers (one tcgen05.ld pass, row max and row sum without any cross-thread traffic), writes P as
// packed bf16 BACK INTO TMEM over the S columns it has consumed, and the P V product reads its A operand from TMEM
// (tcgen05.mma "TS" form) — P never touches shared memory. The O accumulator is rescaled by the same thread, lazily (only when
// the row max grew by more than kRescaleThr), and normalised / stored by it after the last tile.
//
// Semantics are those of attention_tc.cu (W:147-167, V:348-362): mask_mode 1 adds -1e9 in fp32 to keys j <= i, dropout zeroes
// probabilities from the shared (seed, element index) stream and 1/keep is applied with the final normalisation, the row
// statistics (running max, log row sum) go to `stats` for the backward kernels.
#include <math.h>
#include "common.cuh"
#include "ops.cuh"
#include "ptx.cuh"

namespace ts {

int get_tmap(Ctx* ctx, CUtensorMap* out, const void* base, const uint64_t d[4], const uint64_t sbytes[3], uint32_t box0,
             uint32_t box1, bool f32);

namespace {

constexpr int F2_M = 128, F2_N = 128, F2_D = 64;
constexpr int kTile = F2_M * F2_D * 2;          // 16 KB operand tile: [128 x 64] bf16, 128-byte rows, SWIZZLE_128B
constexpr int kKvStages = 4;
constexpr int kThreads2 = 64 + 8 * 32;          // TMA warp, MMA warp, 2 softmax warp-groups
constexpr int kSmem2 = 2 * 2 * kTile /*Q ring: 2 items x (Q0, Q1)*/ + kKvStages * 2 * kTile /*K, V*/ + 256 /*barriers*/ + 1024 /*split mode: (m, l) of group 1*/;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThr = 5.5f;

struct Fwd2Params {
  int B, nh, Tq, Tk, npairs, items;
  int split;   // 1: every item is ONE 128-row query tile whose key tiles alternate between the two softmax groups
  float scale;
  uint32_t drop_thr; float inv_keep; unsigned long long seed;
  const unsigned long long* salt;
  int drop_pitch;
  bf16* o; bf16* o_lo; long long o_ld, o_bs;
  float* stats;
  long long* trace;   // debug (ts_debug_gemm_trace buffer): clock64 stamps of CTA 0's softmax warps 2 and 6, [group][tile 0..15][10]
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int kk) { return ptx::make_smem_desc(tile + kk * 32, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn_b(uint32_t tile, int kk) { return ptx::make_smem_desc(tile + kk * 2048, 8192, 1024); }

template <int T> struct Drop32 {
  static __device__ __forceinline__ void apply(float* v, uint32_t x0, uint32_t thr) {
    v[T] = drop_elem<T>(x0) >= thr ? v[T] : 0.f;
    Drop32<T + 1>::apply(v, x0, thr);
  }
};
template <> struct Drop32<32> {
  static __device__ __forceinline__ void apply(float*, uint32_t, uint32_t) {}
};

struct Item { int b, h, q0; bool has1; };
__device__ __forceinline__ Item decode_item(const Fwd2Params& p, int it) {
  Item w;
  const int pair = it % p.npairs;
  const int bh = it / p.npairs;
  w.h = bh % p.nh; w.b = bh / p.nh;
  w.q0 = pair * (p.split ? F2_M : 2 * F2_M);
  w.has1 = !p.split && w.q0 + F2_M < p.Tq;
  return w;
}

template <int MASK>
__global__ void __launch_bounds__(kThreads2, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const Fwd2Params p, int* watchdog) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0 && watchdog) atomicExch(watchdog, 98);
    return;
  }
  const uint32_t sQ = ptx::smem_u32(smem);                       // [2 items][2 tiles]
  const uint32_t sK = sQ + 4 * kTile, sV = sK + kKvStages * kTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (4 + 2 * kKvStages) * kTile);
  uint64_t* q_full = bars;                   // [2]
  uint64_t* q_empty = bars + 2;              // [2]
  uint64_t* kv_full = bars + 4;              // [kKvStages]
  uint64_t* kv_empty = kv_full + kKvStages;  // [kKvStages]
  uint64_t* s_full = kv_empty + kKvStages;   // [2]  S_w of a tile is in TMEM
  uint64_t* p_full = s_full + 2;             // [2]  P_w written (4 warp arrivals)
  uint64_t* pv_done = p_full + 2;            // [2]  O_w += P_w V retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkv = (p.Tk + F2_N - 1) / F2_N;
  // Split mode (one query tile per item — Tq <= 128, e.g. Whisper's cross-attention 100 x 1500 — and at least two key tiles): the second
  // softmax group would idle, so the key tiles ALTERNATE between the two groups (group g takes tiles g, g + 2, ...), both on query tile 0
  // with their own S / P / O columns and running (max, sum); at the end of the item group 1 hands its partial (m, l, O) to group 0
  // through the unused Q1 tiles of the Q ring and group 0 merges, normalises and stores. Same semantics, half the serial chain.
  const bool split = p.split != 0;   // host: single-query-tile shapes, or where single-tile items fill the SMs better than pairs
  float* sm_ml = reinterpret_cast<float*>(smem + (4 + 2 * kKvStages) * kTile + 256);   // [128][2]

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tm_q); ptx::prefetch_tmap(&tm_k); ptx::prefetch_tmap(&tm_v);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&q_full[s], 1); ptx::mbar_init(&q_empty[s], 1);
      ptx::mbar_init(&s_full[s], 1); ptx::mbar_init(&p_full[s], 4); ptx::mbar_init(&pv_done[s], 1);
    }
    for (int s = 0; s < kKvStages; ++s) { ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ts::pdl_enter();   // prologue above overlaps the previous grid's tail (PDL, common.cuh)
  // TMEM columns: S0 [0,128) (P0 packed bf16 over [0,64)), S1 [128,256) (P1 over [128,192)), O0 [256,320), O1 [320,384)

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t kvc = 0, n = 0;
      bool ok = true;
      for (int it = blockIdx.x; it < p.items && ok; it += gridDim.x, ++n) {
        const Item w = decode_item(p, it);
        const uint32_t qs = n & 1;
        if (!ptx::mbar_wait(&q_empty[qs], ((n >> 1) & 1) ^ 1, watchdog, 41)) break;
        ptx::mbar_expect_tx(&q_full[qs], (w.has1 ? 2 : 1) * kTile);
        ptx::tma_load_4d(sQ + qs * 2 * kTile, &tm_q, &q_full[qs], 0, w.q0, w.h, w.b);
        if (w.has1) ptx::tma_load_4d(sQ + qs * 2 * kTile + kTile, &tm_q, &q_full[qs], 0, w.q0 + F2_M, w.h, w.b);
        for (int j = 0; j < nkv; ++j, ++kvc) {
          const uint32_t s = kvc % kKvStages;
          if (!ptx::mbar_wait(&kv_empty[s], ((kvc / kKvStages) & 1) ^ 1, watchdog, 42)) { ok = false; break; }
          ptx::mbar_expect_tx(&kv_full[s], 2 * kTile);
          ptx::tma_load_4d(sK + s * kTile, &tm_k, &kv_full[s], 0, j * F2_N, w.h, w.b);
          ptx::tma_load_4d(sV + s * kTile, &tm_v, &kv_full[s], 0, j * F2_N, w.h, w.b);
        }
      }
    }
    __syncwarp();
    ts::pdl_tail();   // every operand load of this CTA is in flight: let the next grid's CTAs take the SMs as they free up
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the schedule (uniform control flow and addresses), one elected lane issues =====
    {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(F2_M, F2_N, 0, 0);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(F2_M, F2_D, 0, 1);
      uint32_t kvc = 0, n = 0, pc[2] = {0, 0};   // pc[w]: P_w tiles consumed so far (phase of p_full[w])
      bool ok = true;
      const uint64_t dk0 = desc_kmajor(sK, 0), dv0 = desc_mn_b(sV, 0);
      for (int it = blockIdx.x; it < p.items && ok; it += gridDim.x, ++n) {
        const Item w = decode_item(p, it);
        const uint32_t qs = n & 1;
        const uint32_t q_t[2] = {sQ + qs * 2 * kTile, split ? sQ + qs * 2 * kTile : sQ + qs * 2 * kTile + kTile};
        // descriptors differ only in their 14-bit start-address field (bytes >> 4): one 64-bit add per MMA instead of rebuilding
        // them — the single issuing thread is on the critical path of both softmax groups
        auto issue_s = [&](int g, uint32_t stage) {
          const uint64_t dq = desc_kmajor(q_t[g], 0), dk = dk0 + (uint64_t)(stage * (kTile >> 4));
          if (ptx::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16(tmem + g * 128, dq + (uint64_t)(kk * 2), dk + (uint64_t)(kk * 2), idesc_s, kk > 0 ? 1u : 0u);
            ptx::umma_commit(&s_full[g]);
          }
          __syncwarp();
        };
        auto issue_pv = [&](int g, uint32_t stage, bool acc) {
          const uint64_t dv = dv0 + (uint64_t)(stage * (kTile >> 4));
          if (ptx::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              ptx::umma_f16_ts(tmem + 256 + g * 64, tmem + g * 128 + kk * 8, dv + (uint64_t)(kk * 128), idesc_pv, (acc || kk > 0) ? 1u : 0u);
            ptx::umma_commit(&pv_done[g]);
          }
          __syncwarp();
        };
        auto commit = [&](uint64_t* bar) {
          if (ptx::elect_one()) ptx::umma_commit(bar);
          __syncwarp();
        };
        if (!ptx::mbar_wait(&q_full[qs], (n >> 1) & 1, watchdog, 43)) break;
        if (split) {
          // key tile j belongs to group j & 1; S of tile j + 2 follows P V of tile j of the same group (P aliases S)
          auto kvwait = [&](uint32_t c) { return ptx::mbar_wait(&kv_full[c % kKvStages], (c / kKvStages) & 1, watchdog, 44); };
          if (!kvwait(kvc)) break;
          ptx::tc_fence_after();
          issue_s(0, kvc % kKvStages);
          if (!kvwait(kvc + 1)) break;
          ptx::tc_fence_after();
          issue_s(1, (kvc + 1) % kKvStages);
          for (int j = 0; j < nkv && ok; ++j) {
            const int gj = j & 1;
            const uint32_t st_j = (kvc + j) % kKvStages;
            if (!ptx::mbar_wait(&p_full[gj], pc[gj] & 1, watchdog, 45)) { ok = false; break; }
            ++pc[gj];
            ptx::tc_fence_after();
            issue_pv(gj, st_j, j >= 2);
            if (j + 2 < nkv) {
              if (!kvwait(kvc + j + 2)) { ok = false; break; }
              ptx::tc_fence_after();
              issue_s(gj, (kvc + j + 2) % kKvStages);
            }
            commit(&kv_empty[st_j]);
          }
          kvc += nkv;
          if (ok) commit(&q_empty[qs]);
          continue;
        }
        if (!ptx::mbar_wait(&kv_full[kvc % kKvStages], (kvc / kKvStages) & 1, watchdog, 44)) break;
        ptx::tc_fence_after();
        issue_s(0, kvc % kKvStages);
        if (w.has1) issue_s(1, kvc % kKvStages);
        for (int j = 0; j < nkv && ok; ++j, ++kvc) {
          const uint32_t s = kvc % kKvStages, sn = (kvc + 1) % kKvStages;
          const bool more = j + 1 < nkv;
          if (more && !ptx::mbar_wait(&kv_full[sn], ((kvc + 1) / kKvStages) & 1, watchdog, 44)) { ok = false; break; }
          if (!ptx::mbar_wait(&p_full[0], pc[0] & 1, watchdog, 45)) { ok = false; break; }
          const bool trm = p.trace && blockIdx.x == 0 && pc[0] < 16 && lane == 0;
          if (trm) p.trace[384 + pc[0] * 4 + 0] = clock64();
          ++pc[0];
          ptx::tc_fence_after();
          issue_pv(0, s, j > 0);
          if (trm) p.trace[384 + (pc[0] - 1) * 4 + 1] = clock64();
          if (more) issue_s(0, sn);               // S0 of the next tile overwrites P0 only after P0 V (in-order tensor pipe)
          if (trm) p.trace[384 + (pc[0] - 1) * 4 + 2] = clock64();
          if (w.has1) {
            if (!ptx::mbar_wait(&p_full[1], pc[1] & 1, watchdog, 46)) { ok = false; break; }
            ++pc[1];
            ptx::tc_fence_after();
            issue_pv(1, s, j > 0);
          }
          commit(&kv_empty[s]);                   // K_j / V_j are free once everything issued so far has retired
          if (w.has1 && more) issue_s(1, sn);
        }
        if (ok) commit(&q_empty[qs]);             // every S product of this item has retired: its Q tiles may be overwritten
      }
    }
  } else {
    // ===== softmax warp-groups: g = 0 (warps 2-5, query tile 0) / 1 (warps 6-9, query tile 1); thread = one score row =====
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;                                  // TMEM lane quarter this warp may touch
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem + g * 128 + lane_off, tO = tmem + 256 + g * 64 + lane_off;
    const float c1 = p.scale * kLog2e;
    uint32_t sc = 0, pvc = 0;                                 // S tiles consumed / P V products awaited so far (barrier phases)
    bool ok = true;
    for (int it = blockIdx.x; it < p.items && ok; it += gridDim.x) {
      const Item w = decode_item(p, it);
      if (g == 1 && !w.has1 && !split) continue;
      const int i = w.q0 + (split ? 0 : g * F2_M) + r;        // query index of this thread
      const int myn = split ? (nkv - g + 1) / 2 : nkv;        // key tiles of this group
      const DropKey dkey = make_drop_key(p.drop_thr ? salted_seed(p.seed, p.salt) : p.seed, (unsigned long long)(w.b * p.nh + w.h), p.drop_thr);
      const uint32_t drow = (uint32_t)i * (uint32_t)p.drop_pitch;
      float m_run = -INFINITY, l_run = 0.f;
      for (int jj = 0; jj < myn && ok; ++jj, ++sc) {
        const int j = split ? 2 * jj + g : jj;
        const bool tr = p.trace && blockIdx.x == 0 && (warp == 2 || warp == 6) && lane == 0 && sc < 16;
        long long* trp = p.trace + (g * 16 + (sc & 15)) * 10;
#define F2_STAMP(k) do { if (tr) trp[k] = clock64(); } while (0)
        F2_STAMP(0);
        if (!ptx::mbar_wait(&s_full[g], sc & 1, watchdog, 47)) { ok = false; break; }
        F2_STAMP(1);
        F2_STAMP(2);
        ptx::tc_fence_after();
        uint32_t rg[128];
        {
          uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(rg);
          uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(rg + 32);
          uint32_t(&r2)[32] = *reinterpret_cast<uint32_t(*)[32]>(rg + 64);
          uint32_t(&r3)[32] = *reinterpret_cast<uint32_t(*)[32]>(rg + 96);
          ptx::tmem_ld_32x32(tS, r0); ptx::tmem_ld_32x32(tS + 32, r1); ptx::tmem_ld_32x32(tS + 64, r2); ptx::tmem_ld_32x32(tS + 96, r3);
          ptx::tmem_ld_wait();
        }
        F2_STAMP(3);
        float* sv = reinterpret_cast<float*>(rg);
        const int col0 = j * F2_N;
        constexpr bool plain = MASK == 0;                       // raw accumulators, scale folded into the exponent
        float mx = -INFINITY;
        if (plain) {
          if (col0 + F2_N > p.Tk) {                             // ragged last tile (warp-uniform): keys >= Tk get -inf, p = 0
            const int nvalid = p.Tk - col0;
#pragma unroll
            for (int t = 0; t < 128; ++t) sv[t] = t < nvalid ? sv[t] : -INFINITY;
          }
          float m8[8];                                         // eight independent chains: fmax is not reassociated by the compiler
#pragma unroll
          for (int t = 0; t < 8; ++t) m8[t] = sv[t];
#pragma unroll
          for (int t = 8; t < 128; ++t) m8[t & 7] = fmaxf(m8[t & 7], sv[t]);
          mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
          mx *= p.scale;                                       // scale > 0
        } else {
#pragma unroll
          for (int t = 0; t < 128; ++t) {
            float s = sv[t] * p.scale;
            if (MASK == 1 && col0 + t <= i) s += -1e9f;        // literal fp32 add (absorption is part of the reference's semantics)
            sv[t] = col0 + t < p.Tk ? s : -INFINITY;
            mx = fmaxf(mx, sv[t]);
          }
        }
        if (__any_sync(0xffffffffu, mx > m_run + kRescaleThr)) {
          const float m_new = fmaxf(m_run, mx);
          const float alpha = ex2f((m_run - m_new) * kLog2e);  // 0 on the first tile
          l_run *= alpha;
          if (jj > 0) {   // rescale this row of O once P V of the previous tile has retired (it was issued after our last arrive)
            if (!ptx::mbar_wait(&pv_done[g], (pvc + jj - 1) & 1, watchdog, 48)) { ok = false; break; }
            ptx::tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t ro[16];
              ptx::tmem_ld_32x16(tO + c * 16, ro);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int t = 0; t < 16; ++t) ro[t] = __float_as_uint(__uint_as_float(ro[t]) * alpha);
              ptx::tmem_st_32x16(tO + c * 16, ro);
            }
            ptx::tmem_st_wait();
          }
          m_run = m_new;
        }
        F2_STAMP(4);
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        if (plain) {
          const float nm = -m_run * kLog2e;
#pragma unroll
          for (int t = 0; t < 128; t += 4) {
            sv[t] = ex2f(fmaf(sv[t], c1, nm)); sv[t + 1] = ex2f(fmaf(sv[t + 1], c1, nm));
            sv[t + 2] = ex2f(fmaf(sv[t + 2], c1, nm)); sv[t + 3] = ex2f(fmaf(sv[t + 3], c1, nm));
            l0 += sv[t]; l1 += sv[t + 1]; l2 += sv[t + 2]; l3 += sv[t + 3];
          }
        } else {
#pragma unroll
          for (int t = 0; t < 128; t += 4) {
            sv[t] = ex2f((sv[t] - m_run) * kLog2e); sv[t + 1] = ex2f((sv[t + 1] - m_run) * kLog2e);
            sv[t + 2] = ex2f((sv[t + 2] - m_run) * kLog2e); sv[t + 3] = ex2f((sv[t + 3] - m_run) * kLog2e);
            l0 += sv[t]; l1 += sv[t + 1]; l2 += sv[t + 2]; l3 += sv[t + 3];
          }
        }
        l_run += (l0 + l1) + (l2 + l3);
        F2_STAMP(5);
        if (p.drop_thr) {   // dropped probabilities become 0; the common factor 1/keep is applied to O at the end
#pragma unroll
          for (int c = 0; c < 4; ++c)
            Drop32<0>::apply(sv + 32 * c, drop_chunk_seed(dkey, drow + ((uint32_t)col0 >> 5) + c), dkey.thr);
        }
        F2_STAMP(6);
        // P (bf16 pairs) over the consumed S columns [0, 64) of this row
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) pk[t] = pack2(sv[32 * c + 2 * t], sv[32 * c + 2 * t + 1]);
          ptx::tmem_st_32x16(tS + c * 16, pk);
        }
        ptx::tmem_st_wait();
        F2_STAMP(7);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[g]);
        F2_STAMP(8);
      }
      if (!ok) break;
      // ---- epilogue of this item: O / l -> bf16 -> global, row statistics ----
      const bool tre = p.trace && blockIdx.x == 0 && (warp == 2 || warp == 6) && lane == 0 && pvc < 4u * nkv;
      long long* tep = p.trace + 320 + (g * 4 + (pvc / nkv)) * 8;
#define F2_ESTAMP(k) do { if (tre) tep[k] = clock64(); } while (0)
      F2_ESTAMP(0);
      if (!ptx::mbar_wait(&pv_done[g], (pvc + myn - 1) & 1, watchdog, 49)) { ok = false; break; }
      F2_ESTAMP(1);
      pvc += myn;
      ptx::tc_fence_after();
      // split mode: group 1 parks (m, l, O) of its half of the keys in shared memory, group 0 merges the two partial softmaxes
      float a0 = 1.f, a1 = 0.f;
      float4* part[2] = {reinterpret_cast<float4*>(smem + 1 * kTile), reinterpret_cast<float4*>(smem + 3 * kTile)};   // the Q1 tiles
      if (split) {
        if (g == 1) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t ro[32];
            ptx::tmem_ld_32x32(tO + c * 32, ro);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 8; ++t)    // row r: 8 chunks of 16 bytes, XOR-swizzled against bank conflicts
              part[c][r * 8 + (t ^ (r & 7))] = make_float4(__uint_as_float(ro[4 * t]), __uint_as_float(ro[4 * t + 1]),
                                                           __uint_as_float(ro[4 * t + 2]), __uint_as_float(ro[4 * t + 3]));
          }
          sm_ml[2 * r] = m_run; sm_ml[2 * r + 1] = l_run;
          ptx::tc_fence_before();
          ptx::named_bar_sync(3, 256);      // partial visible to group 0
          ptx::named_bar_sync(4, 256);      // group 0 has read it: the scratch may be rewritten by the next item
          continue;
        }
        ptx::named_bar_sync(3, 256);
        const float m1 = sm_ml[2 * r], l1 = sm_ml[2 * r + 1];
        const float m_all = fmaxf(m_run, m1);
        a0 = ex2f((m_run - m_all) * kLog2e); a1 = ex2f((m1 - m_all) * kLog2e);
        l_run = l_run * a0 + l1 * a1;
        m_run = m_all;
      }
      const float inv = p.inv_keep / l_run;
      const bool live = i < p.Tq;
      const long long ooff = (long long)w.b * p.o_bs + (long long)i * p.o_ld + w.h * F2_D;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ro[32];
        ptx::tmem_ld_32x32(tO + c * 32, ro);
        ptx::tmem_ld_wait();
        F2_ESTAMP(2 + 2 * c);
        if (split) {   // O = O0 a0 + O1 a1 (group 1's partial from shared memory)
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float4 o1 = part[c][r * 8 + (t ^ (r & 7))];
            ro[4 * t] = __float_as_uint(fmaf(__uint_as_float(ro[4 * t]), a0, o1.x * a1));
            ro[4 * t + 1] = __float_as_uint(fmaf(__uint_as_float(ro[4 * t + 1]), a0, o1.y * a1));
            ro[4 * t + 2] = __float_as_uint(fmaf(__uint_as_float(ro[4 * t + 2]), a0, o1.z * a1));
            ro[4 * t + 3] = __float_as_uint(fmaf(__uint_as_float(ro[4 * t + 3]), a0, o1.w * a1));
          }
          if (c == 1) ptx::named_bar_sync(4, 256);   // both halves of the partial have been read
        }
        if (live) {
          float ov[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) ov[t] = __uint_as_float(ro[t]) * inv;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            uint4 u;
            u.x = pack2(ov[8 * t], ov[8 * t + 1]); u.y = pack2(ov[8 * t + 2], ov[8 * t + 3]);
            u.z = pack2(ov[8 * t + 4], ov[8 * t + 5]); u.w = pack2(ov[8 * t + 6], ov[8 * t + 7]);
            reinterpret_cast<uint4*>(p.o + ooff + c * 32)[t] = u;
          }
          if (p.o_lo) {  // bf16 rounding residual, so that backward can form D = rowsum(dO o O) from an (almost) fp32 O
#pragma unroll
            for (int t = 0; t < 32; ++t) ov[t] -= __bfloat162float(__float2bfloat16_rn(ov[t]));
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              uint4 u;
              u.x = pack2(ov[8 * t], ov[8 * t + 1]); u.y = pack2(ov[8 * t + 2], ov[8 * t + 3]);
              u.z = pack2(ov[8 * t + 4], ov[8 * t + 5]); u.w = pack2(ov[8 * t + 6], ov[8 * t + 7]);
              reinterpret_cast<uint4*>(p.o_lo + ooff + c * 32)[t] = u;
            }
          }
        }
        F2_ESTAMP(3 + 2 * c);
      }
      if (live) {
        float* st = p.stats + (((long long)w.b * p.nh + w.h) * p.Tq + i) * 2;
        st[0] = m_run;
        st[1] = logf(l_run);
      }
      // the next item's first P V (accumulate = 0) is issued only after this thread's next p_full arrive: O is free by then
      ptx::tc_fence_before();
      F2_ESTAMP(6);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

}  // namespace

int attn_fwd2(Ctx* ctx, const ts_attn_desc* d, uint32_t drop_thr, float inv_keep, cudaStream_t st) {
  Fwd2Params p;
  memset(&p, 0, sizeof(p));
  p.B = d->batch; p.nh = d->heads; p.Tq = d->tq; p.Tk = d->tk; p.scale = d->scale;
  {
    // work items: pairs of query tiles (K/V tiles shared by the two softmax groups), or single tiles in split mode. Split when there is
    // only one query tile anyway, or when the finer items waste less of the last wave (cost in units of one tile x all its keys on one
    // SM; a split item pays 12-22 % for the merge and the unshared K/V loads: profiles/r02f_selftest_attn_split.log)
    const int nq = cdiv(d->tq, F2_M), nkv = cdiv(d->tk, F2_N), bh = p.B * p.nh, sms = ctx->num_sms;
    const double cost_pair = (double)cdiv((long long)bh * cdiv(nq, 2), sms) * 2.0;
    const double cost_split = (double)cdiv((long long)bh * nq, sms) * 1.25;   // measured 1.12-1.22 per tile against a pair item
    static const int force_split = getenv("TETHYS_ATTN_SPLIT") ? atoi(getenv("TETHYS_ATTN_SPLIT")) : -1;
    p.split = nkv >= 2 && (force_split >= 0 ? force_split != 0 : (nq == 1 || cost_split < cost_pair)) ? 1 : 0;
    p.npairs = p.split ? nq : cdiv(nq, 2);
    p.items = bh * p.npairs;
  }
  p.drop_thr = drop_thr; p.inv_keep = inv_keep; p.seed = d->seed; p.salt = ctx->d_state;
  p.drop_pitch = (d->tk + 31) >> 5;
  p.o = (bf16*)d->o; p.o_lo = (bf16*)d->o_lo; p.o_ld = d->o_ld; p.o_bs = d->o_bs;
  p.stats = d->stats;
  p.trace = reinterpret_cast<long long*>(ctx->gemm_trace);
  auto head_tmap = [&](CUtensorMap* out, const void* base, long long ld, long long bs, int T) {
    const uint64_t dims[4] = {(uint64_t)F2_D, (uint64_t)T, (uint64_t)d->heads, (uint64_t)d->batch};
    const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)F2_D * 2, (uint64_t)(d->batch > 1 ? bs : ld) * 2};
    return get_tmap(ctx, out, base, dims, str, F2_D, 128, false);
  };
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = head_tmap(&tq, d->q, d->q_ld, d->q_bs, d->tq))) return rc;
  if ((rc = head_tmap(&tk, d->k, d->kv_ld, d->kv_bs, d->tk))) return rc;
  if ((rc = head_tmap(&tv, d->v, d->kv_ld, d->kv_bs, d->tk))) return rc;
  static bool attr = false;
  if (!attr) {
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_fwd2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_fwd2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
    attr = true;
  }
  const int grid = p.items < ctx->num_sms ? p.items : ctx->num_sms;
  if (d->mask_mode == 0) ts::launch_k(attn_fwd2_kernel<0>, grid, kThreads2, kSmem2, st, tq, tk, tv, p, ctx->d_watchdog);
  else ts::launch_k(attn_fwd2_kernel<1>, grid, kThreads2, kSmem2, st, tq, tk, tv, p, ctx->d_watchdog);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
