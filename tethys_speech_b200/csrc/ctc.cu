// Connectionist temporal classification loss and its gradient — tf.nn.ctc_loss as Wav2Vec2ForCTC._compute_ctc_loss calls it
// (speech_jobs/whisper_single.py:897-929, the legacy Wav2Vec2 file: dense int labels [B, L], label_length = #(labels > 0),
// logit_length = T, blank_index = 0, time-major logits) — SURVEY §8 f-2.
//
//   extended label sequence l' = (blank, l_1, blank, l_2, ..., l_len, blank), S = 2 len + 1 states
//   y[t, k]   = softmax(logits[t, :])[k]                      (log-space throughout: lp = log y)
//   alpha_t(s) = lp[t, l'_s] + logsumexp(alpha_{t-1}(s), alpha_{t-1}(s-1), alpha_{t-1}(s-2) if l'_s != blank and l'_s != l'_{s-2})
//   loss      = -logsumexp(alpha_{T-1}(S-1), alpha_{T-1}(S-2))
//   beta_t(s)  likewise from the end, emission at t included, so that sum_s alpha_t(s) beta_t(s) / y[t, l'_s] = P for every t
//   d loss / d logits[t, k] = y[t, k] - (1 / P) sum_{s: l'_s = k} alpha_t(s) beta_t(s) / y[t, k]          (Graves et al. 2006, eq. 16)
// One CTA per sample: the T recursion steps are sequential, the S states are spread over the threads; alpha is kept for all t in a
// global workspace [B, T, S] (fp64) because the gradient needs alpha_t and beta_t together. An infinite loss (no valid alignment:
// T < len + repeats) gives loss = +inf and a zero gradient row block; zero_infinity replaces the loss by 0 (WS:920-921).
#include <math.h>
#include "ops.cuh"

namespace ts {

namespace {

// The recursions run in fp64: log alpha reaches ~ -3 T (thousands at T = 750), where an fp32 ulp is ~1e-4 — as a relative error of the
// path probabilities that would be 10x the fp32 parity bar. The kernel is tiny (B x T x S states), fp64 costs nothing here.
__device__ __forceinline__ double lse2(double a, double b) {
  const double m = fmax(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1p(exp(fmin(a, b) - m));
}

template <typename T>
__global__ void __launch_bounds__(256) ctc_kernel(const float* __restrict__ logits, const int* __restrict__ labels, int T_, int V, int L, int blank,
                                                  double* __restrict__ alpha_ws, double* __restrict__ lse_ws, float* __restrict__ loss_out,
                                                  T* __restrict__ dlogits, float grad_scale, int zero_infinity) {
  ts::pdl_enter();
  extern __shared__ double sm[];         // [2][2L+1] recursion rows (fp64), then the extended labels [2L+1] (ints)
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int* lab = labels + (long long)b * L;
  int len = 0;
  for (int i = 0; i < L; ++i) len += lab[i] > 0;       // label_length = reduce_sum(labels > 0) (WS:907); the first `len` entries are used
  const int S = 2 * len + 1, Smax = 2 * L + 1;
  double* row0 = sm;
  double* row1 = sm + Smax;
  int* ext = reinterpret_cast<int*>(sm + 2 * Smax);   // 8-byte rows first: the int tail needs no padding
  for (int s = tid; s < S; s += nt) ext[s] = (s & 1) ? lab[s >> 1] : blank;
  const float* lg = logits + (long long)b * T_ * V;
  double* lse = lse_ws + (long long)b * T_;
  for (int t = tid; t < T_; t += nt) {                 // log-sum-exp of every frame
    double m = -INFINITY;
    for (int k = 0; k < V; ++k) m = fmax(m, (double)lg[(long long)t * V + k]);
    double z = 0.0;
    for (int k = 0; k < V; ++k) z += exp((double)lg[(long long)t * V + k] - m);
    lse[t] = m + log(z);
  }
  __syncthreads();
  double* alpha = alpha_ws + (long long)b * T_ * Smax;
  // ---- forward ----
  for (int s = tid; s < S; s += nt) {
    const double v = s < 2 ? (double)lg[ext[s]] - lse[0] : -INFINITY;
    row0[s] = v;
    alpha[s] = v;
  }
  __syncthreads();
  for (int t = 1; t < T_; ++t) {
    double* prev = (t & 1) ? row0 : row1;
    double* cur = (t & 1) ? row1 : row0;
    for (int s = tid; s < S; s += nt) {
      double a = prev[s];
      if (s >= 1) a = lse2(a, prev[s - 1]);
      if (s >= 2 && ext[s] != blank && ext[s] != ext[s - 2]) a = lse2(a, prev[s - 2]);
      const double v = a == -INFINITY ? -INFINITY : a + (double)lg[(long long)t * V + ext[s]] - lse[t];
      cur[s] = v;
      alpha[(long long)t * Smax + s] = v;
    }
    __syncthreads();
  }
  const double* last = ((T_ - 1) & 1) ? row1 : row0;
  const double logp = S >= 2 ? lse2(last[S - 1], last[S - 2]) : last[S - 1];
  __syncthreads();
  const bool inf = logp == -INFINITY;
  if (tid == 0) loss_out[b] = inf ? (zero_infinity ? 0.f : INFINITY) : (float)(-logp);
  if (!dlogits) return;
  T* dl = dlogits + (long long)b * T_ * V;
  if (inf) {                                           // tf.where(is_inf(loss), 0, loss): no gradient through the replaced entry
    for (long long i = tid; i < (long long)T_ * V; i += nt) dl[i] = from_f<T>(0.f);
    return;
  }
  // ---- backward recursion + gradient, frame by frame from the end ----
  double* beta0 = sm;                                  // both recursion rows are free again
  double* beta1 = sm + Smax;
  for (int t = T_ - 1; t >= 0; --t) {
    double* nxt = (t & 1) ? beta0 : beta1;             // beta_{t+1}
    double* cur = (t & 1) ? beta1 : beta0;
    for (int s = tid; s < S; s += nt) {
      double v;
      if (t == T_ - 1) {
        v = (s >= S - 2) ? (double)lg[(long long)t * V + ext[s]] - lse[t] : -INFINITY;
      } else {
        double a = nxt[s];
        if (s + 1 < S) a = lse2(a, nxt[s + 1]);
        if (s + 2 < S && ext[s + 2] != blank && ext[s + 2] != ext[s]) a = lse2(a, nxt[s + 2]);
        v = a == -INFINITY ? -INFINITY : a + (double)lg[(long long)t * V + ext[s]] - lse[t];
      }
      cur[s] = v;
    }
    __syncthreads();
    // log-occupancy of class k at frame t: logsumexp over the states that emit k of alpha_t(s) + beta_t(s); V is small and the
    // states of one class are few, so a serial scan per class is cheap: thread k owns class k
    for (int k = tid; k < V; k += nt) {
      double o = -INFINITY;
      for (int s = 0; s < S; ++s)
        if (ext[s] == k) o = lse2(o, alpha[(long long)t * Smax + s] + cur[s]);
      const double lp = (double)lg[(long long)t * V + k] - lse[t];
      const double g = exp(lp) - (o == -INFINITY ? 0.0 : exp(o - lp - logp));
      dl[(long long)t * V + k] = from_f<T>((float)(g * (double)grad_scale));
    }
    __syncthreads();
  }
}

}  // namespace

int ctc_loss(Ctx* ctx, int dt, const float* logits, const int* labels, int B, int T_, int V, int L, int blank, float* workspace,
             float* loss_out, void* dlogits, float grad_scale, int zero_infinity, cudaStream_t st) {
  TS_REQUIRE(ctx, logits && labels && loss_out && workspace, TS_EINVAL, "ctc_loss: null pointer");
  TS_REQUIRE(ctx, B > 0 && T_ > 0 && V > 1 && L > 0 && blank >= 0 && blank < V, TS_ESHAPE, "ctc_loss: B=%d T=%d V=%d L=%d blank=%d", B, T_, V, L, blank);
  const int Smax = 2 * L + 1;
  const size_t smem = (size_t)(2 * Smax) * 8 + (size_t)Smax * 4;
  TS_REQUIRE(ctx, smem <= 48 * 1024, TS_ESHAPE, "ctc_loss: label length %d too long for one CTA's shared memory", L);
  TS_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(workspace) & 7) == 0, TS_EINVAL, "ctc_loss: workspace must be 8-byte aligned");
  double* alpha = reinterpret_cast<double*>(workspace);
  double* lse = alpha + (long long)B * T_ * Smax;
  if (dt == TS_F32) ts::launch_k(ctc_kernel<float>, B, 256, smem, st, logits, labels, T_, V, L, blank, alpha, lse, loss_out, (float*)dlogits, grad_scale, zero_infinity);
  else if (dt == TS_BF16) ts::launch_k(ctc_kernel<bf16>, B, 256, smem, st, logits, labels, T_, V, L, blank, alpha, lse, loss_out, (bf16*)dlogits, grad_scale, zero_infinity);
  else return set_err(ctx, TS_EDTYPE, "ctc_loss: dlogits dtype %d", dt);
  TS_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace ts

extern "C" {
int64_t ts_ctc_workspace_floats(int batch, int t, int label_len) { return 2 * ((int64_t)batch * t * (2 * label_len + 1) + (int64_t)batch * t); }
int ts_ctc_loss(ts_ctx* ctx, int grad_dtype, const float* logits, const int32_t* labels, int batch, int t, int vocab, int label_len, int blank,
                float* workspace, float* loss_per_sample, void* dlogits, float grad_scale, int zero_infinity, void* stream) {
  if (!ctx) return TS_EINVAL;
  return ts::ctc_loss(reinterpret_cast<ts::Ctx*>(ctx), grad_dtype, logits, labels, batch, t, vocab, label_len, blank, workspace, loss_per_sample,
                      dlogits, grad_scale, zero_infinity, reinterpret_cast<cudaStream_t>(stream));
}
}
