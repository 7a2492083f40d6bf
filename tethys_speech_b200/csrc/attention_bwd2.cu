// K10 (backward, second generation) — ONE kernel for dQ, dK and dV.
//
// Why: the first-generation backward (attention_tc.cu) ran two kernels that each recomputed S and dP (7 tile products and two
// exponential passes per score tile, 0.10 of the bf16 peak). Here a persistent CTA per SM walks work items (batch, head, 128-key
// tile); per 128-query tile it computes, with keys on the TMEM lanes,
//     S^T = K Q^T, dP^T = V dO^T        (tcgen05, both in TMEM)
//     P^T = exp2(S^T c1 - c0[i]),  dS^T = P^T o (dP^T o Z - D[i]),  PZ = P^T o Z      (one pass, 16 element-wise warps)
//     dV += PZ dO      (PZ stays in TMEM: tcgen05.mma TS form)
//     dK += dS^T Q     (dS^T in smem, K-major A)
//     dQ_i partial = dS K   (the SAME smem tile read as an MN-major A operand)  -> fp32 TMA reduce-add into dq_accum
// i.e. 5 tile products and one exponential per score element. The element-wise warps pull their S^T / dP^T values into
// registers and release the TMEM columns at once, so the tensor pipe computes the next tile's S^T / dP^T underneath the
// exponentials; four more warps drain the dQ partials (TMEM -> swizzled smem -> cp.reduce.async.bulk) and write dK / dV at the
// end of an item, off the element-wise warps' critical path. Two small kernels bracket it: D = rowsum(dO o O) + zeroing of the
// fp32 dQ accumulator before, fp32 -> bf16 of dQ after.
//
// Semantics are those of attention_tc.cu's backward (autodiff transpose of W:147-167 / V:348-362): statistics (row max, log row
// sum) from the forward, dropout keep-masks regenerated from (seed, element), mask_mode 1 = -1e9 added in fp32 to keys j <= i.
#include <math.h>
#include "common.cuh"
#include "ops.cuh"
#include "ptx.cuh"

namespace ts {

int get_tmap(Ctx* ctx, CUtensorMap* out, const void* base, const uint64_t d[4], const uint64_t sbytes[3], uint32_t box0,
             uint32_t box1, bool f32);

namespace {

constexpr int B2_T = 128, B2_D = 64;
constexpr int kTile = B2_T * B2_D * 2;            // 16 KB: [128 x 64] bf16 operand tile (128-byte rows, SWIZZLE_128B)
constexpr int kEw = 16;                           // element-wise warps: warp % 4 = TMEM lane quarter, (warp - 2) / 4 = 32-query chunk
constexpr int kDrain = 4;                         // dQ drain + dK / dV epilogue warps (one per lane quarter)
constexpr int kThreadsB2 = 64 + (kEw + kDrain) * 32;
constexpr int kSmemB2 = 4 * kTile /*K, V x 2 items*/ + 4 * kTile /*Q, dO x 2 tiles*/ + 2 * kTile /*dS^T*/ + 2 * kTile /*dQ staging*/ +
                        2 * 4 * 128 * 4 /*per-query statistics*/ + 256 /*barriers*/;
constexpr float kLog2e = 1.4426950408889634f;

struct Bwd2Params {
  int B, nh, Tq, Tk, nq, nkv, items;
  float scale;
  uint32_t drop_thr; float inv_keep; unsigned long long seed;
  const unsigned long long* salt;
  int drop_pitch;
  const float* stats;   // [B, nh, Tq, 2]
  const float* dsum;    // [B, nh, Tq]  D = rowsum(dO o O)
  bf16 *dk, *dv; long long dkv_ld, dkv_bs;
  long long* trace;   // debug (ts_debug_gemm_trace buffer): clock64 stamps of CTA 0, [role 0 = element-wise warp 2, 1 = MMA warp][tile 0..15][8]
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
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int kk) { return ptx::make_smem_desc(tile + kk * 32, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn_b(uint32_t tile, int kk) { return ptx::make_smem_desc(tile + kk * 2048, 8192, 1024); }
// [128 x 128] bf16 tile stored as two 64-column atoms of 16 KB: K-major A (rows = M) ...
__device__ __forceinline__ uint64_t desc_2atom_kmajor(uint32_t tile, int kk) {
  return ptx::make_smem_desc(tile + (kk >> 2) * kTile + (kk & 3) * 32, 16, 1024);
}
// ... and the same bytes read as an MN-major A (rows = K, the 128 M elements of a row split over the two atoms)
__device__ __forceinline__ uint64_t desc_2atom_mnmajor(uint32_t tile, int kk) { return ptx::make_smem_desc(tile + kk * 2048, kTile, 1024); }
__device__ __forceinline__ uint32_t tile_piece_addr(uint32_t base, int r, int piece) {   // 16-byte piece (8 columns) of row r
  return base + (piece >> 3) * kTile + r * 128 + (((piece & 7) ^ (r & 7)) << 4);
}
__device__ __forceinline__ uint32_t lcg_a_rt(int t) { uint32_t a = 1u; for (int i = 0; i < t; ++i) a *= 1664525u; return a; }
__device__ __forceinline__ uint32_t lcg_c_rt(int t) { uint32_t c = 0u; for (int i = 0; i < t; ++i) c = c * 1664525u + 1013904223u; return c; }

struct ItemB { int b, h, kv0; };
__device__ __forceinline__ ItemB decode_item(const Bwd2Params& p, int it) {
  ItemB w;
  const int j = it % p.nkv;
  const int bh = it / p.nkv;
  w.h = bh % p.nh; w.b = bh / p.nh; w.kv0 = j * B2_T;
  return w;
}

template <int MASK>
__global__ void __launch_bounds__(kThreadsB2, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                 const __grid_constant__ CUtensorMap tm_dq, const Bwd2Params p, int* watchdog) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0 && watchdog) atomicExch(watchdog, 97);
    return;
  }
  const uint32_t sK = ptx::smem_u32(smem), sV = sK + 2 * kTile;          // [2 items]
  const uint32_t sQ = sV + 2 * kTile, sDO = sQ + 2 * kTile;              // [2 tiles]
  const uint32_t sDS = sDO + 2 * kTile;                                  // dS^T [128 keys x 128 queries]
  const uint32_t sDQ = sDS + 2 * kTile;                                  // fp32 staging: 2 boxes [128 rows x 32 cols]
  float* s_stat = reinterpret_cast<float*>(smem + 12 * kTile);           // [2][4: c0, D, m, ll2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 2 * 4 * 128);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* q_full = bars + 4;     // [2]
  uint64_t* q_empty = bars + 6;    // [2]
  uint64_t* sdp_full = bars + 8;   // S^T, dP^T of a tile in TMEM
  uint64_t* sdp_free = bars + 9;   // ... pulled into registers by all element-wise warps
  uint64_t* pds_full = bars + 10;  // PZ (TMEM) and dS^T (smem) of a tile written
  uint64_t* pds_free = bars + 11;  // ... consumed by dV / dK / dQ products
  uint64_t* dq_full = bars + 12;   // dQ partial of a tile in TMEM
  uint64_t* dq_free = bars + 13;   // ... pulled into registers by the drain warps
  uint64_t* acc_full = bars + 14;  // dK, dV of an item complete
  uint64_t* acc_free = bars + 15;  // ... read out
  uint64_t* stat_full = bars + 16; // [2] per-query statistics of a tile staged in smem buffer (tile & 1) by the drain warps. One barrier
                                   // per buffer: consecutive completions of the same barrier are two tiles apart, so a slow waiter can
                                   // never be lapped (a single barrier could complete twice before an element-wise warp tests it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tm_q); ptx::prefetch_tmap(&tm_k); ptx::prefetch_tmap(&tm_v); ptx::prefetch_tmap(&tm_do); ptx::prefetch_tmap(&tm_dq);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&kv_full[s], 1); ptx::mbar_init(&kv_empty[s], 1); ptx::mbar_init(&q_full[s], 1); ptx::mbar_init(&q_empty[s], 1);
    }
    ptx::mbar_init(sdp_full, 1); ptx::mbar_init(sdp_free, kEw);
    ptx::mbar_init(pds_full, kEw); ptx::mbar_init(pds_free, 1);
    ptx::mbar_init(dq_full, 1); ptx::mbar_init(dq_free, kDrain);
    ptx::mbar_init(acc_full, 1); ptx::mbar_init(acc_free, kDrain);
    ptx::mbar_init(&stat_full[0], kDrain); ptx::mbar_init(&stat_full[1], kDrain);
    ptx::fence_barrier_init();
  }
  if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ts::pdl_enter();   // prologue above overlaps the previous grid's tail (PDL, common.cuh)
  const uint32_t tST = tmem, tDPT = tmem + 128, tDV = tmem + 256, tDK = tmem + 320, tDQ = tmem + 384, tPZ = tmem + 448;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t n = 0, tc = 0;
      bool ok = true;
      for (int it = blockIdx.x; it < p.items && ok; it += gridDim.x, ++n) {
        const ItemB w = decode_item(p, it);
        const uint32_t slot = n & 1;
        if (!ptx::mbar_wait(&kv_empty[slot], ((n >> 1) & 1) ^ 1, watchdog, 61)) break;
        ptx::mbar_expect_tx(&kv_full[slot], 2 * kTile);
        ptx::tma_load_4d(sK + slot * kTile, &tm_k, &kv_full[slot], 0, w.kv0, w.h, w.b);
        ptx::tma_load_4d(sV + slot * kTile, &tm_v, &kv_full[slot], 0, w.kv0, w.h, w.b);
        for (int i = 0; i < p.nq; ++i, ++tc) {
          const uint32_t st = tc & 1;
          if (!ptx::mbar_wait(&q_empty[st], ((tc >> 1) & 1) ^ 1, watchdog, 62)) { ok = false; break; }
          ptx::mbar_expect_tx(&q_full[st], 2 * kTile);
          ptx::tma_load_4d(sQ + st * kTile, &tm_q, &q_full[st], 0, i * B2_T, w.h, w.b);
          ptx::tma_load_4d(sDO + st * kTile, &tm_do, &q_full[st], 0, i * B2_T, w.h, w.b);
        }
      }
    }
    __syncwarp();
    ts::pdl_tail();   // every operand load of this CTA is in flight: let the next grid's CTAs take the SMs as they free up
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the schedule, one elected lane issues =====
    constexpr uint32_t idesc_s = ptx::make_idesc_bf16(B2_T, B2_T, 0, 0);       // S^T / dP^T: [128 keys] x [128 queries]
    constexpr uint32_t idesc_acc = ptx::make_idesc_bf16(B2_T, B2_D, 0, 1);     // dV, dK: A K-major (TMEM / smem), B MN-major
    constexpr uint32_t idesc_dq = ptx::make_idesc_bf16(B2_T, B2_D, 1, 1);      // dQ partial: A = dS^T tile read MN-major
    uint32_t n = 0, tc = 0;
    bool ok = true;
    auto commit = [&](uint64_t* bar) {
      if (ptx::elect_one()) ptx::umma_commit(bar);
      __syncwarp();
    };
    for (int it = blockIdx.x; it < p.items && ok; it += gridDim.x, ++n) {
      const uint32_t slot = n & 1;
      const uint32_t kt = sK + slot * kTile, vt = sV + slot * kTile;
      auto issue_sdp = [&](uint32_t t) {
        const uint32_t qt = sQ + (t & 1) * kTile, dot = sDO + (t & 1) * kTile;
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::umma_f16(tST, desc_kmajor(kt, kk), desc_kmajor(qt, kk), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ptx::umma_f16(tDPT, desc_kmajor(vt, kk), desc_kmajor(dot, kk), idesc_s, kk > 0 ? 1u : 0u);
          ptx::umma_commit(sdp_full);
        }
        __syncwarp();
      };
      if (!ptx::mbar_wait(&kv_full[slot], (n >> 1) & 1, watchdog, 63)) break;
      if (!ptx::mbar_wait(&q_full[tc & 1], (tc >> 1) & 1, watchdog, 64)) break;
      ptx::tc_fence_after();
      issue_sdp(tc);
      for (int i = 0; i < p.nq && ok; ++i) {
        const uint32_t t = tc + i;
        const bool trm = p.trace && blockIdx.x == 0 && lane == 0 && t < 16;
        long long* tmp = p.trace + 128 + (t & 15) * 8;
        if (trm) tmp[0] = clock64();
        if (!ptx::mbar_wait(sdp_free, t & 1, watchdog, 65)) { ok = false; break; }
        if (trm) tmp[1] = clock64();
        if (i + 1 < p.nq) {
          if (!ptx::mbar_wait(&q_full[(t + 1) & 1], ((t + 1) >> 1) & 1, watchdog, 64)) { ok = false; break; }
          ptx::tc_fence_after();
          issue_sdp(t + 1);                     // runs underneath the element-wise work on tile t
        }
        if (trm) tmp[2] = clock64();
        if (!ptx::mbar_wait(pds_full, t & 1, watchdog, 66)) { ok = false; break; }
        if (trm) tmp[3] = clock64();
        if (i == 0 && n > 0 && !ptx::mbar_wait(acc_free, (n - 1) & 1, watchdog, 67)) { ok = false; break; }
        if (t > 0 && !ptx::mbar_wait(dq_free, (t - 1) & 1, watchdog, 68)) { ok = false; break; }
        ptx::tc_fence_after();
        const uint32_t qt = sQ + (t & 1) * kTile, dot = sDO + (t & 1) * kTile;
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            ptx::umma_f16_ts(tDV, tPZ + kk * 8, desc_mn_b(dot, kk), idesc_acc, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            ptx::umma_f16(tDK, desc_2atom_kmajor(sDS, kk), desc_mn_b(qt, kk), idesc_acc, (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            ptx::umma_f16(tDQ, desc_2atom_mnmajor(sDS, kk), desc_mn_b(kt, kk), idesc_dq, kk > 0 ? 1u : 0u);
          ptx::umma_commit(pds_free);
          ptx::umma_commit(dq_full);
          ptx::umma_commit(&q_empty[t & 1]);
        }
        __syncwarp();
        if (trm) tmp[4] = clock64();
      }
      if (!ok) break;
      commit(acc_full);
      commit(&kv_empty[slot]);
      tc += p.nq;
    }
  } else if (warp < 2 + kEw) {
    // ===== element-wise warps: thread = one key row (TMEM lane) x 32 query columns =====
    const int qd = warp & 3, chunk = (warp - 2) >> 2;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const float c1 = p.scale * kLog2e;
    const uint32_t a_l = lcg_a_rt(lane + 1), c_l = lcg_c_rt(lane + 1);
    uint32_t n = 0, tc = 0;
    const bool ok = true;
    bool dead = false;   // a bounded wait gave up (pipeline bug): do no more work
    for (int it = blockIdx.x; it < p.items && ok; it += gridDim.x, ++n) {
      const ItemB w = decode_item(p, it);
      const int jrow = w.kv0 + r;                        // key index of this thread
      // key rows >= Tk (TMA zero fill: S^T = dP^T = 0) would give p = exp2(-c0), which overflows where the row maximum is very
      // negative (always under the -1e9 mask) and then poisons dQ through inf * 0: their packed PZ / dS^T words are ANDed away,
      // a branch only the last key tile takes
      const bool ragged_kv = w.kv0 + B2_T > p.Tk;
      const uint32_t rowmask = (jrow < p.Tk) ? 0xffffffffu : 0u;
      const DropKey dkey = make_drop_key(p.drop_thr ? salted_seed(p.seed, p.salt) : p.seed, (unsigned long long)(w.b * p.nh + w.h), p.drop_thr);
      // dropout: this warp's 32 key rows are ONE 32-key chunk (index jrow >> 5) of every query row; lane t hashes the chunk seed of
      // query column t, a shuffle fetches the seed of column c and this lane's jump-ahead constants advance it to its key
      const uint32_t my_chunk = (uint32_t)jrow >> 5;
      for (int i = 0; i < p.nq; ++i) {
        const uint32_t t = tc + i;
        const float* st = s_stat + (t & 1) * 4 * 128;
        const bool tr = p.trace && blockIdx.x == 0 && warp == 2 && lane == 0 && t < 16;
        long long* trp = p.trace + (t & 15) * 8;
#define B2_STAMP(k) do { if (tr) trp[k] = clock64(); } while (0)
        B2_STAMP(0);
        if (dead) continue;
        if (!ptx::mbar_wait(&stat_full[t & 1], (t >> 1) & 1, watchdog, 73)) { dead = true; continue; }   // c0 / D of this tile's 128 queries are in smem
        B2_STAMP(1);
        if (!ptx::mbar_wait(sdp_full, t & 1, watchdog, 69)) { dead = true; continue; }
        B2_STAMP(2);
        ptx::tc_fence_after();
        const int i0 = i * B2_T + chunk * 32;            // first query column of this thread
        const float* sc0 = st + chunk * 32;
        const float* sD = st + 128 + chunk * 32;
        const uint32_t xseed = p.drop_thr ? drop_chunk_seed(dkey, (uint32_t)(i0 + lane) * (uint32_t)p.drop_pitch + my_chunk) : 0u;
        // The 32 columns go through in two halves of 16 (the 80-register budget of a 704-thread CTA holds 16 + 16 inputs, not
        // 32 + 32); the TMEM columns are released after the second load — the next tile's S^T / dP^T products (~500 cycles) still
        // fit underneath the second half's arithmetic and the stores.
        uint32_t pk[16], dk[16];                          // packed bf16 pairs: PZ and dS^T of this thread's 32 columns
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t rs[16], rd[16];
          ptx::tmem_ld_32x16(tST + lane_off + chunk * 32 + hh * 16, rs);
          ptx::tmem_ld_32x16(tDPT + lane_off + chunk * 32 + hh * 16, rd);
          ptx::tmem_ld_wait();
          if (hh == 1) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(sdp_free);   // the tensor pipe may overwrite S^T / dP^T with the next tile's
            B2_STAMP(3);
          }
#pragma unroll
          for (int c4 = 0; c4 < 16; c4 += 4) {
            // per-query statistics: warp-wide broadcast reads, 4 values per LDS
            const float4 qa = *reinterpret_cast<const float4*>(sc0 + hh * 16 + c4), qb = *reinterpret_cast<const float4*>(sD + hh * 16 + c4);
            const float c0v[4] = {qa.x, qa.y, qa.z, qa.w}, Dv[4] = {qb.x, qb.y, qb.z, qb.w};
            float dsv[4], pzv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = c4 + e, col = hh * 16 + c;
              float p0;
              if (MASK == 0) {
                p0 = ex2f(fmaf(__uint_as_float(rs[c]), c1, -c0v[e]));
              } else {
                float s0 = __uint_as_float(rs[c]) * p.scale;
                if (jrow <= i0 + col) s0 += -1e9f;
                p0 = ex2f((s0 - st[256 + chunk * 32 + col]) * kLog2e - st[384 + chunk * 32 + col]);
              }
              if (p.drop_thr) {
                const uint32_t w0 = __shfl_sync(0xffffffffu, xseed, col) * a_l + c_l;
                const bool k0 = w0 >= dkey.thr;
                dsv[e] = p0 * fmaf(__uint_as_float(rd[c]), k0 ? p.inv_keep : 0.f, -Dv[e]);
                pzv[e] = k0 ? p0 : 0.f;
              } else {
                dsv[e] = p0 * (__uint_as_float(rd[c]) - Dv[e]);
                pzv[e] = p0;
              }
            }
            pk[hh * 8 + c4 / 2] = pack2(pzv[0], pzv[1]); pk[hh * 8 + c4 / 2 + 1] = pack2(pzv[2], pzv[3]);
            dk[hh * 8 + c4 / 2] = pack2(dsv[0], dsv[1]); dk[hh * 8 + c4 / 2 + 1] = pack2(dsv[2], dsv[3]);
          }
        }
        if (ragged_kv) {
#pragma unroll
          for (int c = 0; c < 16; ++c) { pk[c] &= rowmask; dk[c] &= rowmask; }
        }
        B2_STAMP(4);
        if (t > 0 && !ptx::mbar_wait(pds_free, (t - 1) & 1, watchdog, 70)) { dead = true; continue; }   // PZ / dS^T of the previous tile consumed
        B2_STAMP(5);
        ptx::tc_fence_after();
        ptx::tmem_st_32x16(tPZ + lane_off + chunk * 16, pk);
#pragma unroll
        for (int c = 0; c < 4; ++c) sts128(tile_piece_addr(sDS, r, chunk * 4 + c), dk[4 * c], dk[4 * c + 1], dk[4 * c + 2], dk[4 * c + 3]);
        ptx::fence_proxy_async_smem();
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(pds_full);
        B2_STAMP(6);
      }
      tc += p.nq;
    }
  } else {
    // ===== drain warps: dQ partial of every tile -> fp32 reduce-add into dq_accum; dK / dV of every item -> global =====
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const bool leader = warp == 2 + kEw && lane == 0;
    uint32_t n = 0, tc = 0;
    bool ok = true, dead = false;
    // per-query statistics (c0 = (m + log l) log2e, D, and m, log2 l for the masked variant) of tile `tile` of item `item` -> smem
    // buffer (tt & 1), one query per thread; the element-wise warps wait on stat_full. Global-load latency lands on these warps,
    // which have slack, instead of on the 16 element-wise warps.
    auto stage_stats = [&](int item, int tile, uint32_t tt) {
      const ItemB ws = decode_item(p, item);
      const long long sbase = ((long long)ws.b * p.nh + ws.h) * p.Tq;
      const int qn = tile * B2_T + r;
      float pm = 0.f, pl = INFINITY, pD = 0.f;           // padding queries: c0 = +inf -> p = 0
      if (qn < p.Tq) { pm = p.stats[(sbase + qn) * 2]; pl = p.stats[(sbase + qn) * 2 + 1] * kLog2e; pD = p.dsum[sbase + qn]; }
      float* st = s_stat + (tt & 1) * 4 * 128;
      st[r] = pm * kLog2e + pl; st[128 + r] = pD; st[256 + r] = pm; st[384 + r] = pl;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&stat_full[tt & 1]);
    };
    if ((int)blockIdx.x < p.items) stage_stats(blockIdx.x, 0, 0);
    for (int it = blockIdx.x; it < p.items && ok; it += gridDim.x, ++n) {
      const ItemB w = decode_item(p, it);
      for (int i = 0; i < p.nq; ++i) {
        const uint32_t t = tc + i;
        // the NEXT tile's statistics (its buffer was last read two tiles ago): next tile of this item, or tile 0 of the next item
        if (i + 1 < p.nq) stage_stats(it, i + 1, t + 1);
        else if (it + (int)gridDim.x < p.items) stage_stats(it + gridDim.x, 0, t + 1);
        uint32_t a0[32], a1[32];
        if (!dead && !ptx::mbar_wait(dq_full, t & 1, watchdog, 71)) dead = true;
        if (dead) {   // keep the two named barriers of this iteration balanced
          asm volatile("bar.sync 6, 128;" ::: "memory");
          asm volatile("bar.sync 6, 128;" ::: "memory");
          continue;
        }
        ptx::tc_fence_after();
        ptx::tmem_ld_32x32(tDQ + lane_off, a0);
        ptx::tmem_ld_32x32(tDQ + lane_off + 32, a1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(dq_free);
        if (leader) ptx::bulk_wait_read<0>();            // the previous tile's reduce has read the staging boxes
        asm volatile("bar.sync 6, 128;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t o = (uint32_t)r * 128 + (uint32_t)((j ^ (r & 7)) << 4);
          sts128(sDQ + o, __float_as_uint(__uint_as_float(a0[4 * j]) * p.scale), __float_as_uint(__uint_as_float(a0[4 * j + 1]) * p.scale),
                 __float_as_uint(__uint_as_float(a0[4 * j + 2]) * p.scale), __float_as_uint(__uint_as_float(a0[4 * j + 3]) * p.scale));
          sts128(sDQ + kTile + o, __float_as_uint(__uint_as_float(a1[4 * j]) * p.scale), __float_as_uint(__uint_as_float(a1[4 * j + 1]) * p.scale),
                 __float_as_uint(__uint_as_float(a1[4 * j + 2]) * p.scale), __float_as_uint(__uint_as_float(a1[4 * j + 3]) * p.scale));
        }
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync 6, 128;" ::: "memory");
        if (leader) {   // rows >= Tq are clipped by the tensor map
          ptx::tma_reduce_add_4d(&tm_dq, sDQ, 0, i * B2_T, w.h, w.b);
          ptx::tma_reduce_add_4d(&tm_dq, sDQ + kTile, 32, i * B2_T, w.h, w.b);
          ptx::bulk_commit();
        }
      }
      if (dead) { ok = false; break; }
      // ---- item epilogue: dV = acc / keep, dK = acc * scale ----
      if (!ptx::mbar_wait(acc_full, n & 1, watchdog, 72)) { ok = false; break; }
      ptx::tc_fence_after();
      const int jrow = w.kv0 + r;
      const long long goff = (long long)w.b * p.dkv_bs + (long long)jrow * p.dkv_ld + w.h * B2_D;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t rg[32];
          ptx::tmem_ld_32x32((which == 0 ? tDV : tDK) + lane_off + c * 32, rg);
          ptx::tmem_ld_wait();
          if (jrow < p.Tk) {
            const float mul = which == 0 ? p.inv_keep : p.scale;
            bf16* dst = (which == 0 ? p.dv : p.dk) + goff + c * 32;
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
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_free);
      tc += p.nq;
    }
    if (leader) ptx::bulk_wait_read<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// D[b, h, i] = sum_d dO * (O + O_lo); zero the fp32 dQ accumulator. Eight consecutive lanes share one (b, i, h) row of 64 elements:
// each loads ONE 16-byte vector of O, O_lo and dO (a warp reads 512 contiguous bytes per tensor and instruction), the partial dot
// products are combined with three shuffles, and every lane zeroes its own 32 bytes of the accumulator row — fully coalesced
// (the one-thread-per-row version touched 32 different 128-byte lines per instruction: 17.5 us for 46 MB).
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ o_lo, const bf16* __restrict__ d_o,
                                                            long long o_ld, long long o_bs, int B, int nh, int Tq, float* __restrict__ dsum,
                                                            float* __restrict__ dq_accum) {
  ts::pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows = (long long)B * Tq * nh;
  const long long r = idx >> 3;                // (b, i, h) row; the 8 lanes of a row are in one warp (blockDim % 8 == 0)
  const int part = (int)(idx & 7);
  const bool ok = r < rows;
  float D = 0.f;
  long long bi = 0;
  int h = 0;
  if (ok) {
    h = (int)(r % nh);
    bi = r / nh;
    const int i = (int)(bi % Tq), b = (int)(bi / Tq);
    const long long off = (long long)b * o_bs + (long long)i * o_ld + h * B2_D + part * 8;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(o + off)), g = __ldg(reinterpret_cast<const uint4*>(d_o + off));
    const uint4 lo = o_lo ? __ldg(reinterpret_cast<const uint4*>(o_lo + off)) : make_uint4(0, 0, 0, 0);
    const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&g);
    const __nv_bfloat162* lh = reinterpret_cast<const __nv_bfloat162*>(&lo);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __bfloat1622float2(ah[e]), y = __bfloat1622float2(gh[e]), z = __bfloat1622float2(lh[e]);
      D = fmaf(x.x + z.x, y.x, D);
      D = fmaf(x.y + z.y, y.y, D);
    }
    float4* z = reinterpret_cast<float4*>(dq_accum + bi * nh * B2_D + h * B2_D + part * 8);
    z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  D += __shfl_xor_sync(0xffffffffu, D, 1);
  D += __shfl_xor_sync(0xffffffffu, D, 2);
  D += __shfl_xor_sync(0xffffffffu, D, 4);
  if (ok && part == 0) {
    const int i = (int)(bi % Tq), b = (int)(bi / Tq);
    dsum[((long long)b * nh + h) * Tq + i] = D;
  }
}

// dq (bf16, caller's strides) = dq_accum (fp32 [B, Tq, nh * 64])
__global__ void __launch_bounds__(256) attn_bwd_dq_store_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long dq_ld,
                                                                long long dq_bs, int B, int nh, int Tq) {
  ts::pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;    // one thread per 8 elements
  const long long per_row = (long long)nh * B2_D / 8;
  if (idx >= (long long)B * Tq * per_row) return;
  const int c8 = (int)(idx % per_row);
  const long long bi = idx / per_row;
  const int i = (int)(bi % Tq), b = (int)(bi / Tq);
  const float4* src = reinterpret_cast<const float4*>(acc + (bi * nh * B2_D) + c8 * 8);
  const float4 x = __ldg(src), y = __ldg(src + 1);
  uint4 u;
  u.x = pack2(x.x, x.y); u.y = pack2(x.z, x.w); u.z = pack2(y.x, y.y); u.w = pack2(y.z, y.w);
  *reinterpret_cast<uint4*>(dq + (long long)b * dq_bs + (long long)i * dq_ld + c8 * 8) = u;
}

}  // namespace

int attn_bwd2(Ctx* ctx, const ts_attn_desc* d, uint32_t drop_thr, float inv_keep, cudaStream_t st) {
  Bwd2Params p;
  memset(&p, 0, sizeof(p));
  p.B = d->batch; p.nh = d->heads; p.Tq = d->tq; p.Tk = d->tk; p.scale = d->scale;
  p.nq = cdiv(d->tq, B2_T); p.nkv = cdiv(d->tk, B2_T);
  p.items = p.B * p.nh * p.nkv;
  p.drop_thr = drop_thr; p.inv_keep = inv_keep; p.seed = d->seed; p.salt = ctx->d_state;
  p.drop_pitch = (d->tk + 31) >> 5;
  p.stats = d->stats; p.dsum = d->dsum;
  p.dk = (bf16*)d->dk; p.dv = (bf16*)d->dv; p.dkv_ld = d->dkv_ld; p.dkv_bs = d->dkv_bs;
  p.trace = reinterpret_cast<long long*>(ctx->gemm_trace);
  auto head_tmap = [&](CUtensorMap* out, const void* base, long long ld, long long bs, int T) {
    const uint64_t dims[4] = {(uint64_t)B2_D, (uint64_t)T, (uint64_t)d->heads, (uint64_t)d->batch};
    const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)B2_D * 2, (uint64_t)(d->batch > 1 ? bs : ld) * 2};
    return get_tmap(ctx, out, base, dims, str, B2_D, 128, false);
  };
  CUtensorMap tq, tk, tv, tdo, tdq;
  int rc;
  if ((rc = head_tmap(&tq, d->q, d->q_ld, d->q_bs, d->tq))) return rc;
  if ((rc = head_tmap(&tk, d->k, d->kv_ld, d->kv_bs, d->tk))) return rc;
  if ((rc = head_tmap(&tv, d->v, d->kv_ld, d->kv_bs, d->tk))) return rc;
  if ((rc = head_tmap(&tdo, d->d_o, d->o_ld, d->o_bs, d->tq))) return rc;
  {
    const long long row = (long long)d->heads * B2_D;
    const uint64_t dims[4] = {(uint64_t)B2_D, (uint64_t)d->tq, (uint64_t)d->heads, (uint64_t)d->batch};
    const uint64_t str[3] = {(uint64_t)row * 4, (uint64_t)B2_D * 4, (uint64_t)d->tq * row * 4};
    if ((rc = get_tmap(ctx, &tdq, d->dq_accum, dims, str, 32, 128, true))) return rc;
  }
  static bool attr = false;
  if (!attr) {
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_bwd2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemB2));
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(attn_bwd2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemB2));
    attr = true;
  }
  const long long rows = (long long)d->batch * d->tq * d->heads;
  ts::launch_k(attn_bwd_prep_kernel, cdiv(rows * 8, 256), 256, 0, st, (const bf16*)d->o, (const bf16*)d->o_lo, (const bf16*)d->d_o, d->o_ld, d->o_bs, d->batch,
                                                       d->heads, d->tq, d->dsum, d->dq_accum);
  TS_LAUNCH_OK(ctx);
  const int grid = p.items < ctx->num_sms ? p.items : ctx->num_sms;
  if (d->mask_mode == 0) ts::launch_k(attn_bwd2_kernel<0>, grid, kThreadsB2, kSmemB2, st, tq, tk, tv, tdo, tdq, p, ctx->d_watchdog);
  else ts::launch_k(attn_bwd2_kernel<1>, grid, kThreadsB2, kSmemB2, st, tq, tk, tv, tdo, tdq, p, ctx->d_watchdog);
  TS_LAUNCH_OK(ctx);
  const long long vec = (long long)d->batch * d->tq * d->heads * B2_D / 8;
  ts::launch_k(attn_bwd_dq_store_kernel, cdiv(vec, 256), 256, 0, st, d->dq_accum, (bf16*)d->dq, d->dq_ld, d->dq_bs, d->batch, d->heads, d->tq);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts
