// K9 — tcgen05 GEMM engine (bf16 in, fp32 accumulate in TMEM, fused epilogue).
//
// One CTA computes one 128 x BN output tile:
//   warp 0 : TMA producer   (cp.async.bulk.tensor 4D boxes -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1 : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16)
//   warps 2-5 : epilogue    (tcgen05.ld TMEM -> registers -> alpha/bias/GELU/residual -> global)
// Both operands may be K-major or MN-major (see include/tethys.h); MN-major operands are loaded as
// 64-wide MN chunks so Dense kernels [in,out], activations for wgrad and V for P.V need no transposes.
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
  long long ldc, ldr, c_bs1, c_bs2, r_bs1, r_bs2, bias_bs1;
  float alpha;
  int act, accumulate;
  int m, n, k, nb1;
  int a_m1, a_m2, b_m1, b_m2;  // 0 => that batch dim is broadcast for the operand (stride 0)
  uint32_t drop_thr; float inv_keep; unsigned long long seed;
};

constexpr int BM = 128;
constexpr int BK = 64;

template <int BN> struct TcCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = (BN < 32) ? 32 : BN;
};

template <typename OutT> __device__ __forceinline__ void store_vec(OutT* dst, const float* v, int n, bool vec_ok);
template <> __device__ __forceinline__ void store_vec<float>(float* dst, const float* v, int n, bool vec_ok) {
  if (vec_ok && n == 32) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    for (int i = 0; i < n; ++i) dst[i] = v[i];
  }
}
template <> __device__ __forceinline__ void store_vec<bf16>(bf16* dst, const float* v, int n, bool vec_ok) {
  if (vec_ok && n == 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * i + 0], v[8 * i + 1]);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
      __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
      u.x = *reinterpret_cast<uint32_t*>(&p0);
      u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2);
      u.w = *reinterpret_cast<uint32_t*>(&p3);
      reinterpret_cast<uint4*>(dst)[i] = u;
    }
  } else {
    for (int i = 0; i < n; ++i) dst[i] = __float2bfloat16_rn(v[i]);
  }
}
template <typename OutT> __device__ __forceinline__ void load_vec(const OutT* src, float* v, int n, bool vec_ok);
template <> __device__ __forceinline__ void load_vec<float>(const float* src, float* v, int n, bool vec_ok) {
  if (vec_ok && n == 32) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 f = reinterpret_cast<const float4*>(src)[i];
      v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
  } else {
    for (int i = 0; i < n; ++i) v[i] = src[i];
  }
}
template <> __device__ __forceinline__ void load_vec<bf16>(const bf16* src, float* v, int n, bool vec_ok) {
  if (vec_ok && n == 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = reinterpret_cast<const uint4*>(src)[i];
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(p[j]);
        v[8 * i + 2 * j] = f.x;
        v[8 * i + 2 * j + 1] = f.y;
      }
    }
  } else {
    for (int i = 0; i < n; ++i) v[i] = __bfloat162float(src[i]);
  }
}

template <int BN, int AMAJ, int BMAJ, typename OutT>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const EpiParams p, int* watchdog) {
  using Cfg = TcCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tmem_full_bar = empty_bar + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int b1 = blockIdx.z % p.nb1;
  const int b2 = blockIdx.z / p.nb1;
  const int nkb = (p.k + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      ptx::prefetch_tmap(&tma_a);
      ptx::prefetch_tmap(&tma_b);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        if (!ptx::mbar_wait(&empty_bar[s], ph ^ 1, watchdog, 1)) break;
        ptx::mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
        const uint32_t sa = ptx::smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
        if (AMAJ == 0) {
          ptx::tma_load_4d(sa, &tma_a, &full_bar[s], kb * BK, m0, b1 * p.a_m1, b2 * p.a_m2);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c)
            ptx::tma_load_4d(sa + c * 8192, &tma_a, &full_bar[s], m0 + c * 64, kb * BK, b1 * p.a_m1, b2 * p.a_m2);
        }
        if (BMAJ == 0) {
          ptx::tma_load_4d(sb, &tma_b, &full_bar[s], kb * BK, n0, b1 * p.b_m1, b2 * p.b_m2);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            ptx::tma_load_4d(sb + c * 8192, &tma_b, &full_bar[s], n0 + c * 64, kb * BK, b1 * p.b_m1, b2 * p.b_m2);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, AMAJ, BMAJ);
      bool ok = true;
      for (int kb = 0; kb < nkb && ok; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        if (!ptx::mbar_wait(&full_bar[s], ph, watchdog, 2)) { ok = false; break; }
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t adesc = (AMAJ == 0) ? ptx::make_smem_desc(sa + kk * 32, 16, 1024)
                                             : ptx::make_smem_desc(sa + kk * 2048, 8192, 1024);
          const uint64_t bdesc = (BMAJ == 0) ? ptx::make_smem_desc(sb + kk * 32, 16, 1024)
                                             : ptx::make_smem_desc(sb + kk * 2048, 8192, 1024);
          ptx::umma_f16(tmem_base, adesc, bdesc, idesc, (kb | kk) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
      }
      ptx::umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool ok = ptx::mbar_wait(tmem_full_bar, 0, watchdog, 3);
    ptx::tc_fence_after();
    if (ok) {
      OutT* crow = reinterpret_cast<OutT*>(p.c) + (long long)b1 * p.c_bs1 + (long long)b2 * p.c_bs2 +
                   (long long)row * p.ldc;
      OutT* prow = p.c_pre ? reinterpret_cast<OutT*>(p.c_pre) + (long long)b1 * p.c_bs1 +
                                 (long long)b2 * p.c_bs2 + (long long)row * p.ldc
                           : nullptr;
      const OutT* rrow = p.res ? reinterpret_cast<const OutT*>(p.res) + (long long)b1 * p.r_bs1 +
                                     (long long)b2 * p.r_bs2 + (long long)row * p.ldr
                               : nullptr;
      const bool c_vec = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
      const bool p_vec = prow && ((reinterpret_cast<uintptr_t>(prow) & 15) == 0);
      const bool r_vec = rrow && ((reinterpret_cast<uintptr_t>(rrow) & 15) == 0);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        ptx::tmem_ld_wait();
        const int col = n0 + c0;
        if (row < p.m && col < p.n) {
          const int nv = min(32, p.n - col);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nv) v[i] += __ldg(p.bias + (long long)b1 * p.bias_bs1 + col + i);
          }
          if (prow) store_vec<OutT>(prow + col, v, nv, p_vec);
          if (p.act == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_f(v[i]);
          }
          if (p.drop_thr) {
            const unsigned long long base = (unsigned long long)((long long)b1 * p.c_bs1 + (long long)b2 * p.c_bs2 +
                                                                 (long long)row * p.ldc + col);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= dropout_scale(p.seed, base + i, p.drop_thr, p.inv_keep);
          }
          if (rrow) {
            float rv[32];
            load_vec<OutT>(rrow + col, rv, nv, r_vec);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nv) v[i] += rv[i];
          }
          if (p.accumulate) {
            float cv[32];
            load_vec<OutT>(crow + col, cv, nv, c_vec);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nv) v[i] += cv[i];
          }
          store_vec<OutT>(crow + col, v, nv, c_vec);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
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
static int get_tmap(Ctx* ctx, CUtensorMap* out, const void* base, const uint64_t d[4], const uint64_t sbytes[3],
                    uint32_t box0, uint32_t box1) {
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
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return 0; }
  cuuint64_t gd[4] = {d[0], d[1], d[2], d[3]};
  cuuint64_t gs[3] = {sbytes[0], sbytes[1], sbytes[2]};
  cuuint32_t bx[4] = {box0, box1, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es,
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

template <int BN, int AMAJ, int BMAJ, typename OutT>
static int launch_tc(Ctx* ctx, const ts_gemm_desc* d, const CUtensorMap& ta, const CUtensorMap& tb,
                     const EpiParams& ep, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  auto kern = gemm_tc_kernel<BN, AMAJ, BMAJ, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_set = true;
  }
  dim3 grid(cdiv(d->m, BM), cdiv(d->n, BN), d->batch1 * d->batch2);
  kern<<<grid, 192, Cfg::kSmem, st>>>(ta, tb, ep, ctx->d_watchdog);
  TS_LAUNCH_OK(ctx);
  return 0;
}

template <int BN, typename OutT>
static int dispatch_major(Ctx* ctx, const ts_gemm_desc* d, const CUtensorMap& ta, const CUtensorMap& tb,
                          const EpiParams& ep, cudaStream_t st) {
  if (d->a_major == 0 && d->b_major == 0) return launch_tc<BN, 0, 0, OutT>(ctx, d, ta, tb, ep, st);
  if (d->a_major == 0 && d->b_major == 1) return launch_tc<BN, 0, 1, OutT>(ctx, d, ta, tb, ep, st);
  if (d->a_major == 1 && d->b_major == 0) return launch_tc<BN, 1, 0, OutT>(ctx, d, ta, tb, ep, st);
  return launch_tc<BN, 1, 1, OutT>(ctx, d, ta, tb, ep, st);
}

int gemm_tc(Ctx* ctx, const ts_gemm_desc* d, cudaStream_t st) {
  TS_REQUIRE(ctx, gemm_tc_supported(d), TS_EUNSUPPORTED,
             "gemm_tc: operands must be bf16 with 16-byte aligned bases/strides");
  TS_REQUIRE(ctx, !(d->accumulate && (d->act || d->residual)), TS_EINVAL, "gemm: accumulate excludes act/residual");
  const int nb1 = d->batch1 > 0 ? d->batch1 : 1, nb2 = d->batch2 > 0 ? d->batch2 : 1;
  // tile-N choice: cover n with the fewest wasted columns; prefer 128.
  int bn = 128;
  if (d->n <= 64) bn = 64;
  else if (d->n % 128 != 0 && d->n % 64 == 0 && d->n < 512) bn = 64;
  else if (d->n >= 1024 && d->n % 256 == 0 && (long long)cdiv(d->m, BM) * (d->n / 256) * nb1 * nb2 >= 2 * ctx->num_sms) bn = 256;

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
    int r = get_tmap(ctx, &tb, d->b, dims, str, 64, d->b_major == 0 ? (uint32_t)bn : BK);
    if (r) return r;
  }
  EpiParams ep;
  ep.c = d->c; ep.c_pre = d->c_preact; ep.res = d->residual; ep.bias = d->bias;
  ep.ldc = d->ldc; ep.ldr = d->ldr; ep.c_bs1 = d->c_bs1; ep.c_bs2 = d->c_bs2; ep.r_bs1 = d->r_bs1; ep.r_bs2 = d->r_bs2; ep.bias_bs1 = d->bias_bs1;
  ep.alpha = d->alpha; ep.act = d->act; ep.accumulate = d->accumulate;
  ep.m = d->m; ep.n = d->n; ep.k = d->k; ep.nb1 = nb1;
  ep.a_m1 = (nb1 > 1 && d->a_bs1 != 0) ? 1 : 0; ep.a_m2 = (nb2 > 1 && d->a_bs2 != 0) ? 1 : 0;
  ep.b_m1 = (nb1 > 1 && d->b_bs1 != 0) ? 1 : 0; ep.b_m2 = (nb2 > 1 && d->b_bs2 != 0) ? 1 : 0;
  ep.drop_thr = 0; ep.inv_keep = 1.f; ep.seed = d->seed;
  if (d->drop > 0.f) {
    double t = (double)d->drop * 4294967296.0;
    ep.drop_thr = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
    ep.inv_keep = 1.f / (1.f - d->drop);
  }
  ts_gemm_desc dd = *d;
  dd.batch1 = nb1; dd.batch2 = nb2;
  if (d->out_dtype == TS_BF16) {
    if (bn == 64) return dispatch_major<64, bf16>(ctx, &dd, ta, tb, ep, st);
    if (bn == 128) return dispatch_major<128, bf16>(ctx, &dd, ta, tb, ep, st);
    return dispatch_major<256, bf16>(ctx, &dd, ta, tb, ep, st);
  } else {
    if (bn == 64) return dispatch_major<64, float>(ctx, &dd, ta, tb, ep, st);
    if (bn == 128) return dispatch_major<128, float>(ctx, &dd, ta, tb, ep, st);
    return dispatch_major<256, float>(ctx, &dd, ta, tb, ep, st);
  }
}

}  // namespace ts
