// Internal C++ API of the kernel library (namespace ts). `dt` is the activation dtype (TS_F32 / TS_BF16);
// parameters, gradients of parameters, statistics and losses are always fp32.
#pragma once
#include <vector>
#include "common.cuh"

namespace ts {

int gemm(Ctx* ctx, const ts_gemm_desc* d, cudaStream_t st);

// ---- norms.cu ----------------------------------------------------------------------------------
int layernorm_fwd(Ctx*, int dt, const void* x, const void* res, const float* gamma, const float* beta, void* y,
                  void* sum_out, float* mean, float* rstd, int rows, int cols, float eps, cudaStream_t);
int layernorm_bwd(Ctx*, int dt, const void* dy, const void* x, const float* gamma, const float* mean,
                  const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int cols,
                  cudaStream_t);
// layernorm_bwd that also emits drop_out = dropout(dx, drop, drop_seed) (flat index row * cols + col) and adds its column sums to
// drop_csum (nullable): the Dropout + bias gradient of the Dense layer that precedes this LayerNorm's input in forward order.
int layernorm_bwd_drop(Ctx*, int dt, const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int cols, void* drop_out, float* drop_csum,
                       float drop, uint64_t drop_seed, cudaStream_t);
// GroupNorm over [B, T(valid rows), C] stored with `rows_per_batch` rows per batch item.
int groupnorm_stats(Ctx*, int dt, const void* x, double* accum /*[B,G,2] scratch*/, float* mean, float* rstd,
                    int B, int T, int C, int G, long long rows_per_batch, float eps, cudaStream_t);
// y[b, out_left + t, :] = gelu(gamma*(x-mean)*rstd+beta); all other rows of y's batch block are zeroed.
int groupnorm_stats_begin(Ctx*, double* accum, int B, int G, cudaStream_t);
int groupnorm_stats_finalize(Ctx*, const double* accum, float* mean, float* rstd, int B, int T, int C, int G, float eps, cudaStream_t);
int groupnorm_gelu_fwd(Ctx*, int dt, const void* x, long long x_rows_per_batch, const float* mean,
                       const float* rstd, const float* gamma, const float* beta, void* y,
                       long long y_rows_per_batch, int y_left, int B, int T, int C, int G, cudaStream_t);
int groupnorm_fwd(Ctx*, int dt, const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta, void* y,
                  int B, int T, int C, int G, cudaStream_t);
// Backward of gelu(GN(x)). The upstream gradient is either dense `da` [B,T,C] (rows_per_batch da_rpb) or
// an im2col gradient `dcol` [B, Tn(+dummy), k*C] of the next strided conv (col2im fused on load):
//   da[b,tau,c] = sum_j dcol[b, (tau+left-j)/s, j*C+c]  for (tau+left-j) % s == 0 and in range.
// pass 1 writes dact = da*gelu'(u) into `dx` and accumulates dgamma/dbeta and per-(b,g) sums;
// pass 2 turns dact into dx in place. dx has dx_rows_per_batch rows per batch; rows >= T are zeroed.
struct Col2imSrc { const void* dcol; long long rows_per_batch; int t_next, k, s, left; };
int groupnorm_gelu_bwd(Ctx*, int dt, const void* da, long long da_rpb, const Col2imSrc* col, const void* x,
                       long long x_rpb, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, void* dx, long long dx_rpb, float* dgamma, float* dbeta,
                       double* accum /*[B,G,2] scratch*/, int B, int T, int C, int G, cudaStream_t, bool skip_pass2 = false);

// ---- elementwise.cu ------------------------------------------------------------------------------
int cast_f32_to_bf16(Ctx*, const float* src, void* dst, long long n, cudaStream_t);
int fill_zero(Ctx*, void* p, long long bytes, cudaStream_t);
// gradient bucket fp32 -> bf16 (times an optional device scalar) and back, for the bf16 all-reduce
int grad_pack_bf16(Ctx*, const float* src, void* dst, long long n, const float* scale_dev, cudaStream_t);
int grad_unpack_bf16(Ctx*, const void* src, float* dst, long long n, cudaStream_t);
// out = gelu(in) (optionally times dropout mask)
int gelu_fwd(Ctx*, int dt, const void* u, void* out, long long n, float drop, uint64_t seed, cudaStream_t);
// du = df * gelu'(u) (* dropout mask scale)
int gelu_bwd(Ctx*, int dt, const void* df, const void* u, void* du, long long n, float drop, uint64_t seed,
             cudaStream_t);
// y = x * dropout_mask (in place allowed)
int dropout_apply(Ctx*, int dt, const void* x, void* y, long long n, float drop, uint64_t seed, cudaStream_t);
// out[n] (+)= column sums of x [rows, cols] (ld). fp32 output, accumulating (atomicAdd).
int colsum_acc(Ctx*, int dt, const void* x, long long ld, int rows, int cols, float* out, cudaStream_t);
// element-wise backward passes fused with the column sums of their output (the bias gradient of the Dense whose dY it is):
// du = df * gelu'(u) * mask / y = x * mask over dense [rows, cols] tensors (cols % 8 == 0); csum[c] += sum_r out[r, c].
int gelu_bwd_colsum(Ctx*, int dt, const void* df, const void* u, void* du, int rows, int cols, float* csum, float drop, uint64_t seed,
                    cudaStream_t);
int dropout_colsum(Ctx*, int dt, const void* x, void* y, int rows, int cols, float* csum, float drop, uint64_t seed, cudaStream_t);
// y = a + b
int add_tensors(Ctx*, int dt, const void* a, const void* b, void* y, long long n, cudaStream_t);
// [B, R, C] -> [B, C, R] style transpose of the two inner dims
int transpose_inner(Ctx*, int dt_in, int dt_out, const void* x, void* y, int B, int R, int C, cudaStream_t);

// ---- conv_fe.cu ----------------------------------------------------------------------------------
// conv0 (Cin = 1): y[b,t,c] = sum_j wave[b, s*t + j - left] * w[j,c];  y has y_rpb rows per batch.
// gn_accum != NULL: also adds the GroupNorm moments of the stored y to gn_accum[B, G, 2] (groupnorm_stats_begin / _finalize).
int conv0_fwd(Ctx*, int dt, const float* wave, const void* w /*act dtype [k,C]*/, void* y, long long y_rpb,
              int B, int N, int T, int C, int k, int s, int left, cudaStream_t, double* gn_accum = nullptr, int G = 0);
// gn_z != NULL: dy is the GELU-backward product of groupnorm_gelu_bwd(..., skip_pass2 = true) and the GroupNorm input gradient
// is formed on load from (dy, gn_z, statistics, pass-1 sums in gn_accum) instead of being written and re-read.
int conv0_wgrad(Ctx*, int dt, const float* wave, const void* dy, long long dy_rpb, float* dw /*[k,C] +=*/,
                int B, int N, int T, int C, int k, int s, int left, cudaStream_t, const void* gn_z = nullptr, long long gn_z_rpb = 0,
                const float* gn_mean = nullptr, const float* gn_rstd = nullptr, const float* gn_gamma = nullptr,
                const double* gn_accum = nullptr, int G = 0);
// group-major repack for the grouped positional conv: x [B,T,C] -> xg [G][B][T+K-1][C/G] with
// `left` zero rows before and K-1-left after each block.
int posconv_pack(Ctx*, int dt, const void* x, void* xg, int B, int T, int C, int G, int K, int left,
                 cudaStream_t);
// flipped/transposed kernel for dgrad: wt[g][K-1-j][o][c] = w[j][c][g*cpg+o]   (w is [K, cpg, C])
int posconv_flip_weight(Ctx*, int dt, const void* w, void* wt, int K, int C, int G, cudaStream_t);

// ---- attention.cu --------------------------------------------------------------------------------
// in-place row softmax of scores [rows_total = nbatch*Tq, ld] (valid cols Tk) with scale and mask mode:
//   mask 0: none; 1: whisper decoder "anti-causal" additive mask (-1e9 on j<=i, fp32 absorption, App. C-1).
int softmax_fwd(Ctx*, int dt, void* s, long long ld, int nbatch, int Tq, int Tk, float scale, int mask_mode,
                float drop, uint64_t seed, void* p_drop /*optional dropped probs out, same layout*/, cudaStream_t);
// dS = scale * P * (dP - rowsum(dP*P)) in place over dP; with dropout the mask is re-applied to dP first.
int softmax_bwd(Ctx*, int dt, const void* p, void* dp, long long ld, int nbatch, int Tq, int Tk, float scale,
                float drop, uint64_t seed, cudaStream_t);

// ---- attention_tc.cu: fused tcgen05 attention (bf16, head_dim 64) ----------------------------------
bool attn_tc_supported(const ts_attn_desc* d);
int attn_fwd(Ctx*, const ts_attn_desc* d, cudaStream_t);
int attn_bwd(Ctx*, const ts_attn_desc* d, cudaStream_t);

// ---- logmel.cu: K1 log-mel front end (W:739-766) ------------------------------------------------------
// wave [batch, n_samples] fp32 (batch stride wave_bs) -> log-mel [batch, F, 80] (mel_major = 0) or [batch, 80, F] (1)
int logmel(Ctx*, const float* wave, long long wave_bs, int batch, int n_samples, void* out, int out_dtype, int mel_major,
           int sample_rate, int n_mels, int n_fft, int hop, cudaStream_t);

// ---- vq.cu / contrastive.cu ----------------------------------------------------------------------
// hard VQ (V:604-660): z [M, G*D] (act dtype), codebook fp32 [G,V,D]; writes q [M, G*D], idx int64 [G,M],
// hist int32 [G,V] (+=), then perplexity (fp32 scalar) via vq_perplexity.
int vq_fwd(Ctx*, int dt, const void* z, const float* codebook, void* q, long long* idx, int* hist, int M, int G,
           int V, int D, cudaStream_t);
int vq_perplexity(Ctx*, const int* hist, float* perplexity, int M, int G, int V, cudaStream_t);
// dcodebook[g, idx[g,m], :] += dq[m, g*D:(g+1)*D]
int vq_bwd(Ctx*, int dt, const void* dq, const long long* idx, float* dcodebook, int M, int G, int V, int D,
           cudaStream_t);
// contrastive loss on the all-pairs similarity S [B,T,T] (fp32, ld): logits = {S[t,t], S[t,neg[b,t,k]]}/temp,
// loss = mean CE(label 0). Writes dS (act dtype, ld_ds; zero except the K+1 touched entries, duplicates
// accumulate) = dloss/dS * grad_scale, the logits [B,T,K+1] (optional) and accumulates the loss sum.
int contrastive_fwd_bwd(Ctx*, int dt, const float* S, long long ld, const int* neg, long long neg_bs,
                        long long neg_ts, void* dS, long long ld_ds, float* logits, float* loss_sum, int B, int T,
                        int K, float temp, float grad_scale, cudaStream_t);

// ---- task_heads.cu: CTC / sequence-classification heads on the Wav2Vec2 trunk (V:940-1070) ----------
// row-wise sparse softmax CE on fp32 logits [R, ld] (labels NULL = class 0 everywhere, V:997-1000): accumulates the loss
// sum and writes dlogits (act dtype, [R, ld_d]; nullable) = (softmax - onehot) * grad_scale.
int ce_rows_fwd_bwd(Ctx*, int dt, const float* logits, long long ld, const int* labels, void* dlogits, long long ld_d,
                    float* loss_sum, int R, int V, float grad_scale, cudaStream_t);
int mean_pool_fwd(Ctx*, int dt, const void* x /*[B,T,C]*/, void* y /*[B,C]*/, int B, int T, int C, cudaStream_t);
int mean_pool_bwd(Ctx*, int dt, const void* dy /*[B,C]*/, void* dx /*[B,T,C]*/, int B, int T, int C, cudaStream_t);
int tanh_drop_fwd(Ctx*, int dt, const void* x, void* y /*tanh(x)*/, void* y_drop /*tanh(x)*mask*/, long long n, float drop,
                  uint64_t seed, cudaStream_t);
int tanh_drop_bwd(Ctx*, int dt, const void* dy, const void* y, void* dx, long long n, float drop, uint64_t seed, cudaStream_t);

// ---- loss_ce.cu (whisper) ------------------------------------------------------------------------
// shifted sparse softmax CE over logits [B,S,ldv] (valid V): rows s<S-1 with target labels[b,s+1];
// writes dlogits (act dtype, in place over logits allowed; row S-1 and pad cols get 0) scaled by
// grad_scale/(B*(S-1)), accumulates loss sum.
int ce_fwd_bwd(Ctx*, int dt, const void* logits, void* dlogits, long long ldv, const int* labels, float* loss_sum,
               int B, int S, int V, float grad_scale, cudaStream_t);
// embedding gather + positional encoding (+ dropout): out[b,s,:] = drop(table[ids[b,s],:] + pe[s,:]); the ids are the
// labels shifted right with the start token (W:559-563). embed_bwd scatter-adds drop(dout) into the fp32 table grad.
int embed_fwd(Ctx*, int dt, const void* table, const int* labels, long long label_ld /*ints per batch row*/, const float* pe, void* out, int B, int S, int D,
              int start_token, float drop, uint64_t seed, cudaStream_t);
int embed_bwd(Ctx*, int dt, const void* dout, const int* labels, float* dtable, int B, int S, int D,
              int start_token, float drop, uint64_t seed, cudaStream_t);
// y[b,t,:] = drop(x[b,t,:] + pe[t,:]); x has rpb_in rows per batch, y is dense [B,T,D]
int add_pe_rows(Ctx*, int dt, const void* x, long long rpb_in, const float* pe, void* y, int B, int T, int D, float drop,
                uint64_t seed, cudaStream_t);
// du = da * gelu'(u) on [B, rpb, C] blocks (rows >= T zeroed); da dense (da_rpb) or col2im of a strided conv's dcol
int gelu_bwd_rows(Ctx*, int dt, const void* da, long long da_rpb, const Col2imSrc* col, const void* u, void* du,
                  long long rpb, int B, int T, int C, cudaStream_t);
int zero_rows(Ctx*, int dt, void* x, long long rpb, int row_from, int B, int C, cudaStream_t);

// ---- ctc.cu: tf.nn.ctc_loss (dense labels, blank_index given) per sample + d loss / d logits * grad_scale (dlogits nullable) ----
int ctc_loss(Ctx*, int dt, const float* logits, const int* labels, int B, int T, int V, int L, int blank, float* workspace,
             float* loss_out, void* dlogits, float grad_scale, int zero_infinity, cudaStream_t);

// ---- optim.cu ------------------------------------------------------------------------------------
// A parameter "segment": rows x cols block with row stride ld inside the flat arena (dense: rows = 1).
struct Segment { long long offset; int rows, cols; long long ld; };
// One block's share of a segment (<= 64K elements); built once per optimizer object by build_work_items.
struct WorkItem { long long start; long long ld; int rows, cols, seg, pad; };
std::vector<WorkItem> build_work_items(const std::vector<Segment>& segs);
// per-segment sum of squares of grads -> sumsq[nseg] (fp32, overwritten)
int grad_sumsq(Ctx*, const void* grads, int grad_dt /*TS_F32 | TS_BF16*/, const WorkItem* d_items, int nitems, int nseg, float* sumsq, cudaStream_t);
// scales[0] = global clip scale from sum(sumsq) (clip_by_global_norm, V:1243) or 1
int global_clip_scale(Ctx*, const float* sumsq, int nseg, float clip, float* scale_out, float* norm_out,
                      cudaStream_t);
// per-segment clip scale after multiplying by pre_scale[0]: s_i = clip/max(||g_i||*pre, clip)   (V:1274)
// Adam (Keras-2.10 legacy, App. A-12) over every segment with g' = g * pre_scale * seg_scale.
struct AdamArgs {
  float lr, beta1, beta2, eps; int step;     // step t >= 1
  float clipnorm;                            // <=0: no per-variable clip
  const float* pre_scale;                    // device scalar or NULL (=1)
  const float* sumsq;                        // per segment sum of squares of the (unscaled) grads, or NULL
};
int adam_step(Ctx*, float* params, const void* grads, int grad_dt /*TS_F32 | TS_BF16*/, float* m, float* v, void* params_bf16 /*or NULL*/,
              const WorkItem* d_items, int nitems, const AdamArgs& a, cudaStream_t);
int scale_inplace(Ctx*, float* x, long long n, const float* scale_dev, float scale_host, cudaStream_t);

}  // namespace ts
