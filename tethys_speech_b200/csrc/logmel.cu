// K1 — log-mel front end: replaces extract_fbank_features (W:739-766)
//   tf.signal.stft(x, 400, 160, fft_length=400) -> |.|^2 -> linear_to_mel_weight_matrix(80, 201, 16000, 0, 8000) -> ln(. + 1e-6)
// in ONE kernel: the waveform is read once (6 % overlap between blocks), the [frames, 201] spectra never leave the SM.
//
// A block owns 24 consecutive frames of one sample (= 12 complex 400-point FFTs: two real frames ride on the real and
// imaginary part of one transform and are separated by conjugate symmetry). The 400-point FFT is a 20 x 20 four-step:
//   X[k1 + 20 k2] = sum_n2 W20^{n2 k2} * ( W400^{n2 k1} * sum_n1 x[20 n1 + n2] W20^{n1 k1} )
// with each length-20 DFT done by one thread entirely in registers (4 x 5 Cooley-Tukey: 5 radix-4 + 12 twiddles +
// 4 radix-5 butterflies, ~550 flop) — ~12 kflop per frame instead of 320 kflop for the DFT-as-GEMM formulation, so the
// kernel is bound by the 4 N + 4 * 80 * F bytes per sample it must move, not by arithmetic. fp32 throughout (the
// reference is fp32); window / twiddle / mel tables are computed in double on the host once per process.
#include <math.h>
#include <mutex>
#include "common.cuh"
#include "ops.cuh"

namespace ts {
namespace {

constexpr int LM_NFFT = 400, LM_HOP = 160, LM_BINS = 201, LM_MELS = 80;
constexpr int LM_FPB = 24;                                   // frames per block
constexpr int LM_PAIRS = LM_FPB / 2;
constexpr int LM_WAVE = (LM_FPB - 1) * LM_HOP + LM_NFFT;     // 4080 samples staged per block
constexpr int LM_ROW = 21;                                   // padded row (float2) of the 20 x 20 intermediate
constexpr int LM_MAXW = 16;                                  // widest mel triangle in bins (13 for the reference config)

struct LogmelTables {
  float hann[LM_NFFT];
  float2 w400[LM_NFFT];          // e^{-2 pi i m / 400}
  int mel_start[LM_MELS];
  int mel_cnt[LM_MELS];
  float mel_w[LM_MELS][LM_MAXW];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// a * (-i) and a * (+i)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }
__device__ __forceinline__ float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }

// forward DFT of length 20, in place: x[n] -> X[k], n = 5a + b, k = c + 4d
__device__ __forceinline__ void dft20(float2 (&x)[20]) {
  constexpr float wr[5][4] = {{1.0f, 1.0f, 1.0f, 1.0f},
                              {1.0f, 0.9510565162951535f, 0.8090169943749475f, 0.5877852522924731f},
                              {1.0f, 0.8090169943749475f, 0.30901699437494745f, -0.30901699437494734f},
                              {1.0f, 0.5877852522924731f, -0.30901699437494734f, -0.9510565162951535f},
                              {1.0f, 0.30901699437494745f, -0.8090169943749473f, -0.8090169943749476f}};
  constexpr float wi[5][4] = {{0.0f, 0.0f, 0.0f, 0.0f},
                              {0.0f, -0.3090169943749474f, -0.5877852522924731f, -0.8090169943749475f},
                              {0.0f, -0.5877852522924731f, -0.9510565162951535f, -0.9510565162951536f},
                              {0.0f, -0.8090169943749475f, -0.9510565162951536f, -0.3090169943749475f},
                              {0.0f, -0.9510565162951535f, -0.5877852522924732f, 0.587785252292473f}};
  float2 t[5][4];
#pragma unroll
  for (int b = 0; b < 5; ++b) {  // radix-4 over a, then the W20^{bc} twiddle
    const float2 u0 = x[b], u1 = x[5 + b], u2 = x[10 + b], u3 = x[15 + b];
    const float2 s02 = cadd(u0, u2), d02 = csub(u0, u2), s13 = cadd(u1, u3), d13 = csub(u1, u3);
    t[b][0] = cadd(s02, s13);
    t[b][1] = cadd(d02, mul_mi(d13));
    t[b][2] = csub(s02, s13);
    t[b][3] = cadd(d02, mul_pi(d13));
    if (b > 0) {
#pragma unroll
      for (int c = 1; c < 4; ++c) t[b][c] = cmul(t[b][c], make_float2(wr[b][c], wi[b][c]));
    }
  }
  constexpr float c1 = 0.30901699437494745f, c2 = -0.8090169943749473f, s1 = 0.9510565162951535f, s2 = 0.5877852522924732f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {  // radix-5 over b -> X[c + 4d]
    const float2 y0 = t[0][c], y1 = t[1][c], y2 = t[2][c], y3 = t[3][c], y4 = t[4][c];
    const float2 t1 = cadd(y1, y4), t2 = cadd(y2, y3), t3 = csub(y1, y4), t4 = csub(y2, y3);
    const float2 m1 = make_float2(y0.x + c1 * t1.x + c2 * t2.x, y0.y + c1 * t1.y + c2 * t2.y);
    const float2 m2 = make_float2(y0.x + c2 * t1.x + c1 * t2.x, y0.y + c2 * t1.y + c1 * t2.y);
    const float2 q1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    const float2 q2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    x[c] = make_float2(y0.x + t1.x + t2.x, y0.y + t1.y + t2.y);
    x[c + 4] = cadd(m1, mul_mi(q1));
    x[c + 16] = cadd(m1, mul_pi(q1));
    x[c + 8] = cadd(m2, mul_mi(q2));
    x[c + 12] = cadd(m2, mul_pi(q2));
  }
}

constexpr int kLogmelSmem = LM_WAVE * 4 + LM_PAIRS * 20 * LM_ROW * 8 /*buf*/;   // 56.6 KB: four blocks per SM

template <typename TO>
__global__ void __launch_bounds__(256, 4) logmel_kernel(const float* __restrict__ wave, long long wave_bs, int n_samples,
                                                     const LogmelTables* __restrict__ tab, TO* __restrict__ out, int n_frames,
                                                     int mel_major /*0: [B,F,80]  1: [B,80,F]*/) {
  ts::pdl_enter();
  extern __shared__ __align__(16) unsigned char lm_smem[];
  float* s_wave = reinterpret_cast<float*>(lm_smem);
  float2* s_buf = reinterpret_cast<float2*>(s_wave + LM_WAVE);
  const float* g_hann = tab->hann;        // window / twiddle tables stay in global memory: 4.8 KB, L1-resident
  const float2* g_w = tab->w400;
  const int b = blockIdx.y, f0 = blockIdx.x * LM_FPB;
  const int tid = threadIdx.x;
  // ---- stage the waveform chunk ----------------------------------------------------------------------------------
  const float* src = wave + (long long)b * wave_bs + (long long)f0 * LM_HOP;
  const int avail = n_samples - f0 * LM_HOP;  // samples left in this sample from the chunk start
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = tid * 4; i < LM_WAVE; i += 1024) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i + 3 < avail) v = __ldg(reinterpret_cast<const float4*>(src + i));
      else {
        if (i < avail) v.x = src[i];
        if (i + 1 < avail) v.y = src[i + 1];
        if (i + 2 < avail) v.z = src[i + 2];
      }
      *reinterpret_cast<float4*>(s_wave + i) = v;
    }
  } else {
    for (int i = tid; i < LM_WAVE; i += 256) s_wave[i] = i < avail ? src[i] : 0.f;
  }
  __syncthreads();
  const int p = tid / 20, j = tid % 20;  // (frame pair, column / row of the 20 x 20 decomposition)
  const bool fft_thread = tid < LM_PAIRS * 20;
  float2 z[20];
  // ---- step 1: windowed pair -> DFT-20 over n1 -> twiddle W400^{n2 k1} -----------------------------------------
  if (fft_thread) {
    const float* fa = s_wave + (2 * p) * LM_HOP;
    const float* fb = fa + LM_HOP;
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
      const int n = 20 * n1 + j;
      const float w = __ldg(g_hann + n);
      z[n1] = make_float2(w * fa[n], w * fb[n]);
    }
    dft20(z);
#pragma unroll
    for (int k1 = 0; k1 < 20; ++k1) s_buf[(p * 20 + k1) * LM_ROW + j] = cmul(z[k1], __ldg(g_w + j * k1));
  }
  __syncthreads();
  // ---- step 2: DFT-20 over n2 for row k1 = j -> X[k1 + 20 k2] --------------------------------------------------
  if (fft_thread) {
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) z[n2] = s_buf[(p * 20 + j) * LM_ROW + n2];
    dft20(z);
  }
  __syncthreads();
  float2* s_x = s_buf;  // spectrum of pair p at s_x[p * 400 + k] (fits: 400 <= 20 * LM_ROW)
  if (fft_thread) {
#pragma unroll
    for (int k2 = 0; k2 < 20; ++k2) s_x[p * 20 * LM_ROW + j + 20 * k2] = z[k2];
  }
  __syncthreads();
  // ---- power (conjugate symmetry separates the two real frames of a pair) + sparse mel filter bank + log -----------------
  for (int it = tid; it < LM_FPB * LM_MELS; it += 256) {
    int fr, m;
    if (mel_major) { m = it / LM_FPB; fr = it % LM_FPB; } else { fr = it / LM_MELS; m = it % LM_MELS; }
    const int f = f0 + fr;
    if (f >= n_frames) continue;
    const float2* X = s_x + (fr >> 1) * 20 * LM_ROW;
    const bool second = fr & 1;
    const int st = tab->mel_start[m], cnt = tab->mel_cnt[m];
    float acc = 0.f;
    for (int q = 0; q < cnt; ++q) {
      const int k = st + q;                       // 1 <= k <= 200
      const float2 a = X[k], c = X[LM_NFFT - k];
      // frame A = (X[k] + conj(X[N-k])) / 2, frame B = (X[k] - conj(X[N-k])) / (2i)
      const float re = second ? a.x - c.x : a.x + c.x, im = second ? a.y + c.y : a.y - c.y;
      acc = fmaf(__ldg(&tab->mel_w[m][q]), 0.25f * (re * re + im * im), acc);
    }
    const float v = __logf(acc + 1e-6f);
    const long long o = mel_major ? ((long long)b * LM_MELS + m) * n_frames + f : ((long long)b * n_frames + f) * LM_MELS + m;
    out[o] = from_f<TO>(v);
  }
}

LogmelTables* g_tables[16] = {nullptr};
std::mutex g_tab_mu;

static int get_tables(Ctx* ctx, LogmelTables** out) {
  std::lock_guard<std::mutex> lk(g_tab_mu);
  const int dev = ctx->device & 15;
  if (!g_tables[dev]) {
    LogmelTables* h = new LogmelTables();
    memset(h, 0, sizeof(*h));
    const double pi = 3.14159265358979323846;
    for (int n = 0; n < LM_NFFT; ++n) {
      h->hann[n] = (float)(0.5 - 0.5 * cos(2.0 * pi * n / LM_NFFT));  // periodic Hann (tf.signal.hann_window default)
      h->w400[n] = make_float2((float)cos(2.0 * pi * n / LM_NFFT), (float)(-sin(2.0 * pi * n / LM_NFFT)));
    }
    // tf.signal.linear_to_mel_weight_matrix(80, 201, 16000, 0, 8000): HTK mel, triangles in the mel domain, DC bin zeroed
    auto mel = [](double f) { return 1127.0 * log1p(f / 700.0); };
    double edges[LM_MELS + 2];
    for (int i = 0; i < LM_MELS + 2; ++i) edges[i] = mel(0.0) + (mel(8000.0) - mel(0.0)) * i / (LM_MELS + 1);
    for (int m = 0; m < LM_MELS; ++m) {
      int start = -1, cnt = 0;
      for (int k = 1; k < LM_BINS; ++k) {
        const double fm = mel(8000.0 * k / (LM_BINS - 1));
        const double lo = (fm - edges[m]) / (edges[m + 1] - edges[m]), up = (edges[m + 2] - fm) / (edges[m + 2] - edges[m + 1]);
        const double w = fmax(0.0, fmin(lo, up));
        if (w > 0.0) {
          if (start < 0) start = k;
          if (k - start >= LM_MAXW) { delete h; return set_err(ctx, TS_EUNSUPPORTED, "logmel: mel filter wider than %d bins", LM_MAXW); }
          h->mel_w[m][k - start] = (float)w;
          cnt = k - start + 1;
        }
      }
      h->mel_start[m] = start < 0 ? 0 : start;
      h->mel_cnt[m] = cnt;
    }
    LogmelTables* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(LogmelTables));
    if (e == cudaSuccess) e = cudaMemcpy(d, h, sizeof(LogmelTables), cudaMemcpyHostToDevice);
    delete h;
    if (e != cudaSuccess) return set_err(ctx, TS_ECUDA, "logmel tables: %s", cudaGetErrorString(e));
    g_tables[dev] = d;
  }
  *out = g_tables[dev];
  return 0;
}

}  // namespace

int logmel(Ctx* ctx, const float* wave, long long wave_bs, int batch, int n_samples, void* out, int out_dtype, int mel_major,
           int sample_rate, int n_mels, int n_fft, int hop, cudaStream_t st) {
  TS_REQUIRE(ctx, sample_rate == 16000 && n_mels == LM_MELS && n_fft == LM_NFFT && hop == LM_HOP, TS_EUNSUPPORTED,
             "logmel: only the reference configuration (16 kHz, 80 mels, n_fft 400, hop 160) is implemented");
  TS_REQUIRE(ctx, batch > 0 && wave && out, TS_EINVAL, "logmel: bad arguments");
  const int nf = n_samples < LM_NFFT ? 0 : 1 + (n_samples - LM_NFFT) / LM_HOP;
  if (nf == 0) return 0;  // tf.signal.stft(pad_end=False) on a short signal: zero frames
  LogmelTables* tab;
  int rc = get_tables(ctx, &tab);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(logmel_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLogmelSmem));
    TS_CUDA_OK(ctx, cudaFuncSetAttribute(logmel_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLogmelSmem));
    attr = true;
  }
  dim3 grid(cdiv(nf, LM_FPB), batch);
  if (out_dtype == TS_F32) ts::launch_k(logmel_kernel<float>, grid, 256, kLogmelSmem, st, wave, wave_bs, n_samples, tab, (float*)out, nf, mel_major);
  else if (out_dtype == TS_BF16) ts::launch_k(logmel_kernel<bf16>, grid, 256, kLogmelSmem, st, wave, wave_bs, n_samples, tab, (bf16*)out, nf, mel_major);
  else return set_err(ctx, TS_EDTYPE, "logmel: output dtype %d", out_dtype);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts

extern "C" {
int ts_logmel_num_frames(int n_samples) { return n_samples < 400 ? 0 : 1 + (n_samples - 400) / 160; }
int ts_logmel(ts_ctx* ctx, const float* wave, int64_t wave_batch_stride, int batch, int n_samples, void* out, int out_dtype,
              int mel_major, void* stream) {
  if (!ctx) return TS_EINVAL;
  return ts::logmel(reinterpret_cast<ts::Ctx*>(ctx), wave, wave_batch_stride, batch, n_samples, out, out_dtype, mel_major, 16000, 80, 400,
                    160, reinterpret_cast<cudaStream_t>(stream));
}
}
