// Wav2Vec2 pre-training step as a native program over pre-bound arenas: forward (V:768-825, V:841-863), the
// contrastive + diversity loss (V:865-905, V:1208-1220) and the hand-derived backward of every op, calling the
// kernel library. All GEMM-shaped work (strided convs as window-GEMMs without im2col, the grouped positional
// conv as per-group window-GEMMs, Dense layers, attention products, the all-pairs contrastive similarities)
// goes through ts::gemm (tcgen05 in bf16 mode, fp32 CUDA cores in parity mode).
#include <stdlib.h>
#include "program.cuh"

namespace ts {

struct W2VLayerOff { long long qkv_w, qkv_b, o_w, o_b, ln1_g, ln1_b, fc1_w, fc1_b, fc2_w, fc2_b, ln2_g, ln2_b; };
struct W2VLayerBuf {
  void *h_in, *x1, *qkv, *P, *ctx, *ctx_lo, *h_mid, *x2, *u, *f;
  float *ln1_mean, *ln1_rstd, *ln2_mean, *ln2_rstd;
};

struct W2V {
  Ctx* ctx = nullptr;
  ts_w2v_config cfg;
  int prec = TS_F32, esz = 4;
  bool fused_attn = false;  // bf16 + head_dim 64: tcgen05 flash kernels instead of GEMM/softmax/GEMM
  ParamTable pt;
  // parameter offsets
  long long conv_w[8], conv_g[8], conv_b[8];
  long long pos_w, pos_b, fe_ln_g, fe_ln_b, fp_w, fp_b, fp_ln_g, fp_ln_b;
  std::vector<W2VLayerOff> L;
  long long cb, qp_w, qp_b, ph_w, ph_b, ph_ln_g, ph_ln_b, pq_w, pq_b, pq_ln_g, pq_ln_b;
  long long lm_w = -1, lm_b = -1, cp_w = -1, cp_b = -1, cl_w = -1, cl_b = -1;   // task heads (cfg.head 1 / 2)
  std::vector<long long> stage_end;  // arena offset below which grads are final after backward stage s
  // bound memory
  float *P = nullptr, *G = nullptr;
  void* P16 = nullptr;
  char* ws = nullptr;
  long long ws_bytes = 0;
  // geometry of the current plan
  int B = 0, N = 0, nconv = 0, T = 0, Tp = 0, M = 0;
  int Tc[8], padl[8], padr[8], Rq[8];
  long long rpb_c[8], rpb_a[8];
  int a_left[8];
  // buffers
  void *c[8], *a[8];
  float *gn_mean[8], *gn_rstd[8];
  double* gn_accum;
  bool no_fused_gn = getenv("TETHYS_NO_FUSED_GN") && atoi(getenv("TETHYS_NO_FUSED_GN")) != 0;   // A/B switch: separate statistics pass
  void *hg, *possum, *ef, *fp_out, *hs, *z, *qfeat, *ph_lin, *ps, *pq_lin, *pq, *dS, *wt_flip;
  float *S, *logits;
  // task heads: dropped encoder output, fp32 logits + their gradient; pooled / projected states of the classifier
  void *hd_in, *d_hlogits, *pooled, *cp_pre, *cp_act, *cp_drop, *d_cp, *d_pooled;
  float* hlogits;
  int head_rows = 0, head_cols = 0;
  float *fe_mean, *fe_rstd, *fp_mean, *fp_rstd, *ph_mean, *ph_rstd, *pq_mean, *pq_rstd;
  long long* code_idx;
  int* hist;
  float* scalars;  // [0] loss, [1] contrastive, [2] perplexity, [3] loss_sum accumulator
  std::vector<W2VLayerBuf> LB;
  void* enc_out;
  // backward scratch
  void *g_a, *g_b, *g_t, *g_x, *g_f, *g_ctx, *g_qkv, *g_P, *g_Pd, *g_dqacc, *g_small1, *g_small2, *g_dcol, *g_dc, *g_dyg;
  // state of the last forward
  uint64_t seed = 0;
  int training = 1;
  float loss_div = 1.f;
  const float* wave = nullptr;
  bool planned = false, fwd_done = false;

  const void* W(long long off) const { return prec == TS_BF16 ? (const void*)((const bf16*)P16 + off) : (const void*)(P + off); }
  size_t E(long long n) const { return (size_t)n * esz; }
  float drop(float r) const { return training ? r : 0.f; }
};

static void build_params(W2V* m) {
  const ts_w2v_config& c = m->cfg;
  ParamTable& pt = m->pt;
  const int H = c.hidden, F = c.ffn, C = c.conv_dim[c.n_conv - 1], D = c.cv_dim, Pj = c.proj_dim;
  // arena order = order in which gradients become final during backward (bucketed all-reduce overlap)
  if (c.head == 1) {          // Wav2Vec2ForCTC: lm_head (V:950); project_hid / project_q are never built
    m->lm_w = pt.add("lm_head.kernel", {H, c.vocab_size});
    m->lm_b = pt.add("lm_head.bias", {c.vocab_size});
  } else if (c.head == 2) {   // Wav2Vec2ForSequenceClassification: classifier_proj + classifier (V:1013-1015)
    m->cl_w = pt.add("classifier.kernel", {c.classifier_proj, c.num_labels});
    m->cl_b = pt.add("classifier.bias", {c.num_labels});
    m->cp_w = pt.add("classifier_proj.kernel", {H, c.classifier_proj});
    m->cp_b = pt.add("classifier_proj.bias", {c.classifier_proj});
  }
  if (c.head == 0) {
  m->pq_w = pt.add("project_q.dense.kernel", {D, Pj});
  m->pq_b = pt.add("project_q.dense.bias", {Pj});
  m->pq_ln_g = pt.add("project_q.layer_norm.gamma", {Pj});
  m->pq_ln_b = pt.add("project_q.layer_norm.beta", {Pj});
  m->ph_w = pt.add("project_hid.dense.kernel", {H, Pj});
  m->ph_b = pt.add("project_hid.dense.bias", {Pj});
  m->ph_ln_g = pt.add("project_hid.layer_norm.gamma", {Pj});
  m->ph_ln_b = pt.add("project_hid.layer_norm.beta", {Pj});
  }
  m->cb = pt.add("quantizer.codevectors", {c.cv_groups, c.cv_per_group, D / c.cv_groups});
  m->qp_w = pt.add("quantizer.projection.kernel", {H, D});
  m->qp_b = pt.add("quantizer.projection.bias", {D});
  m->stage_end.push_back(pt.n);
  m->L.resize(c.layers);
  for (int l = c.layers - 1; l >= 0; --l) {
    const std::string p = "encoder.layers." + std::to_string(l) + ".";
    W2VLayerOff& o = m->L[l];
    o.fc2_w = pt.add(p + "feed_forward.output_dense.kernel", {F, H});
    o.fc2_b = pt.add(p + "feed_forward.output_dense.bias", {H});
    o.fc1_w = pt.add(p + "feed_forward.intermediate_dense.kernel", {H, F});
    o.fc1_b = pt.add(p + "feed_forward.intermediate_dense.bias", {F});
    o.ln2_g = pt.add(p + "feed_forward_layer_norm.gamma", {H});
    o.ln2_b = pt.add(p + "feed_forward_layer_norm.beta", {H});
    o.o_w = pt.add(p + "attention.out_proj.kernel", {H, H});
    o.o_b = pt.add(p + "attention.out_proj.bias", {H});
    o.qkv_w = pt.add_fused({p + "attention.q_proj.kernel", p + "attention.k_proj.kernel", p + "attention.v_proj.kernel"}, H, H, 3 * H);
    o.qkv_b = pt.add_fused({p + "attention.q_proj.bias", p + "attention.k_proj.bias", p + "attention.v_proj.bias"}, 1, H, 3 * H);
    o.ln1_g = pt.add(p + "attention_layer_norm.gamma", {H});
    o.ln1_b = pt.add(p + "attention_layer_norm.beta", {H});
    m->stage_end.push_back(pt.n);
  }
  m->fp_ln_g = pt.add("feature_projection_layer_norm.gamma", {H});
  m->fp_ln_b = pt.add("feature_projection_layer_norm.beta", {H});
  m->fp_w = pt.add("feature_projection.kernel", {C, H});
  m->fp_b = pt.add("feature_projection.bias", {H});
  m->fe_ln_g = pt.add("fe.layer_norm.gamma", {C});
  m->fe_ln_b = pt.add("fe.layer_norm.beta", {C});
  m->pos_w = pt.add("fe.pos_conv.kernel", {c.pos_kernel, C / c.pos_groups, C});
  m->pos_b = pt.add("fe.pos_conv.bias", {C});
  for (int i = c.n_conv - 1; i >= 0; --i) {
    const std::string p = "fe.conv" + std::to_string(i) + ".";
    const int cin = i == 0 ? 1 : c.conv_dim[i - 1];
    m->conv_g[i] = pt.add(p + "gn.gamma", {c.conv_dim[i]});
    m->conv_b[i] = pt.add(p + "gn.beta", {c.conv_dim[i]});
    m->conv_w[i] = pt.add(p + "kernel", {c.conv_kernel[i], cin, c.conv_dim[i]});
  }
  pt.n = (pt.n + 63) & ~63ll;
  m->stage_end.push_back(pt.n);
}

// Lay out every buffer of a (B, N) step in the workspace. dry run (base == nullptr) only measures.
static int plan(W2V* m, int B, int N, Bump& bp) {
  const ts_w2v_config& c = m->cfg;
  Ctx* ctx = m->ctx;
  m->B = B; m->N = N; m->nconv = c.n_conv;
  const int n = c.n_conv, G = c.pos_groups;
  int t = N;
  for (int i = 0; i < n; ++i) {
    int to, l, r;
    same_pad(t, c.conv_kernel[i], c.conv_stride[i], &to, &l, &r);
    m->Tc[i] = to; m->padl[i] = l; m->padr[i] = r;
    t = to;
  }
  TS_REQUIRE(ctx, t >= 2, TS_ESHAPE, "w2v: %d samples give %d frames (need >= 2)", N, t);
  m->T = t; m->Tp = (t + 7) & ~7; m->M = B * t;
  // rows per batch: a[i-1] feeds conv i through windows of stride s_i; R = s_i * Rq_i
  for (int i = 0; i < n; ++i) {
    if (i + 1 < n) {
      const int s = c.conv_stride[i + 1], k = c.conv_kernel[i + 1];
      long long need = (long long)m->padl[i + 1] + m->Tc[i] + m->padr[i + 1];
      long long R = std::max<long long>((long long)s * (m->Tc[i + 1] + 1), need + (k > s ? k - s : 0));
      R = (R + s - 1) / s * s;
      m->rpb_a[i] = R;
      m->Rq[i + 1] = (int)(R / s);
      m->a_left[i] = m->padl[i + 1];
    } else {
      m->rpb_a[i] = m->Tc[i];
      m->a_left[i] = 0;
    }
  }
  m->Rq[0] = m->Tc[0];
  for (int i = 0; i < n; ++i) m->rpb_c[i] = (i == 0) ? m->Tc[0] : m->Rq[i];
  const int C = c.conv_dim[n - 1], H = c.hidden, F = c.ffn, D = c.cv_dim, Pj = c.proj_dim, T = m->T, Tp = m->Tp, M = m->M;
  const int nh = c.heads;
  TS_REQUIRE(ctx, H % nh == 0 && (H / nh) % 8 == 0, TS_ESHAPE, "w2v: hidden %d / heads %d", H, nh);
  for (int i = 0; i < n; ++i) TS_REQUIRE(ctx, c.conv_dim[i] % 8 == 0, TS_ESHAPE, "w2v: conv_dim must be a multiple of 8");
  for (int i = 0; i < n; ++i) {
    const long long slack = (i + 1 < n) ? (long long)c.conv_kernel[i + 1] * c.conv_dim[i] : 0;
    m->c[i] = bp.get(m->E((long long)B * m->rpb_c[i] * c.conv_dim[i]));
    m->a[i] = bp.get(m->E((long long)B * m->rpb_a[i] * c.conv_dim[i] + slack));
    m->gn_mean[i] = (float*)bp.get(sizeof(float) * B * G);
    m->gn_rstd[i] = (float*)bp.get(sizeof(float) * B * G);
  }
  m->gn_accum = (double*)bp.get(sizeof(double) * B * G * 2);
  const int K = c.pos_kernel, cpg = C / G, Rp = T + K - 1;
  m->hg = bp.get(m->E((long long)G * B * Rp * cpg + (long long)K * cpg));
  m->wt_flip = bp.get(m->E((long long)K * cpg * C));
  m->possum = bp.get(m->E((long long)M * C));
  m->ef = bp.get(m->E((long long)M * C));
  m->fe_mean = (float*)bp.get(4 * M); m->fe_rstd = (float*)bp.get(4 * M);
  m->fp_out = bp.get(m->E((long long)M * H));
  m->hs = bp.get(m->E((long long)M * H));
  m->fp_mean = (float*)bp.get(4 * M); m->fp_rstd = (float*)bp.get(4 * M);
  m->z = bp.get(m->E((long long)M * D));
  m->qfeat = bp.get(m->E((long long)M * D));
  m->code_idx = (long long*)bp.get(8ll * c.cv_groups * M);
  m->hist = (int*)bp.get(4ll * c.cv_groups * c.cv_per_group);
  m->scalars = (float*)bp.get(64);
  m->LB.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    W2VLayerBuf& b = m->LB[l];
    if (l == 0) b.h_in = m->hs;
    b.x1 = bp.get(m->E((long long)M * H));
    b.qkv = bp.get(m->E((long long)M * 3 * H));
    b.P = m->fused_attn ? bp.get(8ll * B * nh * T) : bp.get(m->E((long long)B * nh * T * Tp));   // fused: [B,nh,T,2] row statistics only
    b.ctx = bp.get(m->E((long long)M * H));
    b.ctx_lo = m->fused_attn ? bp.get(m->E((long long)M * H)) : nullptr;
    b.h_mid = bp.get(m->E((long long)M * H));
    b.x2 = bp.get(m->E((long long)M * H));
    b.u = bp.get(m->E((long long)M * F));
    b.f = bp.get(m->E((long long)M * F));
    b.ln1_mean = (float*)bp.get(4 * M); b.ln1_rstd = (float*)bp.get(4 * M);
    b.ln2_mean = (float*)bp.get(4 * M); b.ln2_rstd = (float*)bp.get(4 * M);
    void* h_out = bp.get(m->E((long long)M * H));
    if (l + 1 < c.layers) m->LB[l + 1].h_in = h_out; else m->enc_out = h_out;
  }
  if (c.layers == 0) m->enc_out = m->hs;
  if (c.head == 0) {
    m->ph_lin = bp.get(m->E((long long)M * Pj)); m->ps = bp.get(m->E((long long)M * Pj));
    m->pq_lin = bp.get(m->E((long long)M * Pj)); m->pq = bp.get(m->E((long long)M * Pj));
    m->ph_mean = (float*)bp.get(4 * M); m->ph_rstd = (float*)bp.get(4 * M);
    m->pq_mean = (float*)bp.get(4 * M); m->pq_rstd = (float*)bp.get(4 * M);
    m->S = (float*)bp.get(4ll * B * T * Tp);
    m->dS = bp.get(m->E((long long)B * T * Tp));
    m->logits = (float*)bp.get(4ll * M * (c.num_negatives + 1));
  } else if (c.head == 1) {
    m->head_rows = M; m->head_cols = c.vocab_size;
    m->hd_in = bp.get(m->E((long long)M * H));
    m->hlogits = (float*)bp.get(4ll * M * c.vocab_size);
    m->d_hlogits = bp.get(m->E((long long)M * c.vocab_size));
  } else {
    const int Pc = c.classifier_proj;
    m->head_rows = B; m->head_cols = c.num_labels;
    m->pooled = bp.get(m->E((long long)B * H));
    m->cp_pre = bp.get(m->E((long long)B * Pc)); m->cp_act = bp.get(m->E((long long)B * Pc)); m->cp_drop = bp.get(m->E((long long)B * Pc));
    m->hlogits = (float*)bp.get(4ll * B * c.num_labels);
    m->d_hlogits = bp.get(m->E((long long)B * c.num_labels));
    m->d_cp = bp.get(m->E((long long)B * Pc));
    m->d_pooled = bp.get(m->E((long long)B * H));
  }
  // backward scratch
  m->g_a = bp.get(m->E((long long)M * std::max(H, C)));
  m->g_b = bp.get(m->E((long long)M * std::max(H, C)));
  m->g_t = bp.get(m->E((long long)M * std::max(H, C)));
  m->g_x = bp.get(m->E((long long)M * std::max(H, C)));
  m->g_f = bp.get(m->E((long long)M * F));
  m->g_ctx = bp.get(m->E((long long)M * H));
  m->g_qkv = bp.get(m->E((long long)M * 3 * H));
  if (m->fused_attn) { m->g_P = bp.get(4ll * B * nh * T); m->g_Pd = nullptr; m->g_dqacc = bp.get(4ll * B * T * H); }   // fused: D = rowsum(dO o O) scratch + the fp32 dQ accumulator
  else { m->g_P = bp.get(m->E((long long)B * nh * T * Tp)); m->g_Pd = bp.get(m->E((long long)B * nh * T * Tp)); m->g_dqacc = nullptr; }
  m->g_small1 = bp.get(m->E((long long)M * std::max(Pj, D)));
  m->g_small2 = bp.get(m->E((long long)M * std::max(Pj, D)));
  m->g_dyg = bp.get(m->E((long long)G * B * Rp * cpg + (long long)K * cpg));
  long long dcol_max = 0, dc_max = 0;
  for (int i = 0; i < n; ++i) {
    dc_max = std::max(dc_max, (long long)B * m->rpb_c[i] * c.conv_dim[i]);
    if (i >= 1) dcol_max = std::max(dcol_max, (long long)B * m->Rq[i] * c.conv_kernel[i] * c.conv_dim[i - 1]);
  }
  m->g_dcol = bp.get(m->E(dcol_max));
  m->g_dc = bp.get(m->E(dc_max));
  return 0;
}

__global__ void w2v_finalize_scalars(float* s, float inv_rows, float div_w) {
  ts::pdl_enter();
  const float closs = s[3] * inv_rows;
  float loss = closs + div_w * (-s[2]);
  if (isnan(loss)) loss = 0.f;   // tf.where(is_nan(loss), 0, loss) — V:1228
  s[1] = closs;
  s[0] = loss;
}

static int w2v_forward(W2V* m, const float* wave, const int* neg, long long neg_bs, long long neg_ts, cudaStream_t st,
                       bool features_only = false, const int* labels = nullptr, bool with_loss = true) {
  Ctx* ctx = m->ctx;
  const ts_w2v_config& c = m->cfg;
  const int dt = m->prec, B = m->B, n = m->nconv, G = c.pos_groups;
  const int T = m->T, Tp = m->Tp, M = m->M, H = c.hidden, F = c.ffn, nh = c.heads, hd = H / nh;
  const int C = c.conv_dim[n - 1], D = c.cv_dim, Pj = c.proj_dim;
  const uint64_t seed = m->seed;
  m->wave = wave;
  TS_TRY(fill_zero(ctx, m->scalars, 64, st));
  // ---- conv feature encoder (V:283-288) --------------------------------------------------------------
  for (int i = 0; i < n; ++i) {
    const int Ci = c.conv_dim[i];
    // GroupNorm moments (V:167-176) are taken by the producer of the conv output where it can: conv0's store loop and the tcgen05
    // GEMM epilogue (channels per group a multiple of 32); otherwise (fp32 parity engine, tiny test configs) by a pass over c[i]
    const bool fused_stats = !m->no_fused_gn && (i == 0 || (dt == TS_BF16 && (Ci / G) % 32 == 0));
    if (fused_stats) TS_TRY(groupnorm_stats_begin(ctx, m->gn_accum, B, G, st));
    if (i == 0) {
      TS_TRY(conv0_fwd(ctx, dt, wave, m->P + m->conv_w[0], m->c[0], m->rpb_c[0], B, m->N, m->Tc[0], Ci, c.conv_kernel[0],
                       c.conv_stride[0], m->padl[0], st, fused_stats ? m->gn_accum : nullptr, G));
    } else {
      const int Cin = c.conv_dim[i - 1], k = c.conv_kernel[i], s = c.conv_stride[i];
      GemmB g(dt, dt);
      g.A(m->a[i - 1], 0, (long long)s * Cin).B(m->W(m->conv_w[i]), 1, Ci).C(m->c[i], Ci).mnk(B * m->Rq[i], Ci, k * Cin);
      if (fused_stats) g.gn_stats(m->gn_accum, m->Rq[i], m->Tc[i], G);
      TS_TRY(g.run(ctx, st));
    }
    if (fused_stats) TS_TRY(groupnorm_stats_finalize(ctx, m->gn_accum, m->gn_mean[i], m->gn_rstd[i], B, m->Tc[i], Ci, G, 1e-5f, st));
    else TS_TRY(groupnorm_stats(ctx, dt, m->c[i], m->gn_accum, m->gn_mean[i], m->gn_rstd[i], B, m->Tc[i], Ci, G, m->rpb_c[i], 1e-5f, st));
    TS_TRY(groupnorm_gelu_fwd(ctx, dt, m->c[i], m->rpb_c[i], m->gn_mean[i], m->gn_rstd[i], m->P + m->conv_g[i],
                              m->P + m->conv_b[i], m->a[i], m->rpb_a[i], m->a_left[i], B, m->Tc[i], Ci, G, st));
    if (i + 1 < n)  // finite slack behind the last batch block (window reads of the dummy rows)
      TS_TRY(fill_zero(ctx, (char*)m->a[i] + m->E((long long)B * m->rpb_a[i] * Ci), m->E((long long)c.conv_kernel[i + 1] * Ci), st));
  }
  // ---- grouped positional conv + residual + LayerNorm (V:291-296) --------------------------------------
  {
    const int K = c.pos_kernel, cpg = C / G, Rp = T + K - 1, left = (K - 1) / 2;
    TS_TRY(posconv_pack(ctx, dt, m->a[n - 1], m->hg, B, T, C, G, K, left, st));
    TS_TRY(fill_zero(ctx, (char*)m->hg + m->E((long long)G * B * Rp * cpg), m->E((long long)K * cpg), st));
    TS_TRY(GemmB(dt, dt).A(m->hg, 0, cpg).astride((long long)B * Rp * cpg, (long long)Rp * cpg)
               .B(m->W(m->pos_w), 1, C).bstride(cpg, 0)
               .C(m->possum, C).cstride(cpg, (long long)T * C)
               .res(m->a[n - 1], C, cpg, (long long)T * C)
               .bias(m->P + m->pos_b, cpg).mnk(T, cpg, K * cpg).batch(G, B).run(ctx, st));
    TS_TRY(layernorm_fwd(ctx, dt, m->possum, nullptr, m->P + m->fe_ln_g, m->P + m->fe_ln_b, m->ef, nullptr, m->fe_mean,
                         m->fe_rstd, M, C, c.ln_eps, st));
    if (m->drop(c.hidden_dropout) > 0) TS_TRY(dropout_apply(ctx, dt, m->ef, m->ef, (long long)M * C, c.hidden_dropout, site_seed(seed, 1), st));
  }
  if (features_only) return 0;   // Wav2Vec2FeatureExtractor.call only (V:283-298): front-end microbench / feature export
  // ---- feature projection (V:777-779) -------------------------------------------------------------------
  TS_TRY(GemmB(dt, dt).A(m->ef, 0, C).B(m->W(m->fp_w), 1, H).C(m->fp_out, H).bias(m->P + m->fp_b).mnk(M, H, C).run(ctx, st));
  TS_TRY(layernorm_fwd(ctx, dt, m->fp_out, nullptr, m->P + m->fp_ln_g, m->P + m->fp_ln_b, m->hs, nullptr, m->fp_mean,
                       m->fp_rstd, M, H, c.ln_eps, st));
  if (m->drop(c.hidden_dropout) > 0) TS_TRY(dropout_apply(ctx, dt, m->hs, m->hs, (long long)M * H, c.hidden_dropout, site_seed(seed, 2), st));
  // ---- quantiser on the projected states (V:784-789, V:581-667); the task heads run it in training mode only --------
  if (c.head == 0 || with_loss) {
    const int Gq = c.cv_groups, V = c.cv_per_group, Dg = D / Gq;
    TS_TRY(GemmB(dt, dt).A(m->hs, 0, H).B(m->W(m->qp_w), 1, D).C(m->z, D).bias(m->P + m->qp_b).mnk(M, D, H).run(ctx, st));
    TS_TRY(fill_zero(ctx, m->hist, 4ll * Gq * V, st));
    TS_TRY(vq_fwd(ctx, dt, m->z, m->P + m->cb, m->qfeat, m->code_idx, m->hist, M, Gq, V, Dg, st));
    TS_TRY(vq_perplexity(ctx, m->hist, m->scalars + 2, M, Gq, V, st));
  }
  // ---- transformer encoder, pre-LN (V:419-439) -----------------------------------------------------------
  for (int l = 0; l < c.layers; ++l) {
    const W2VLayerOff& o = m->L[l];
    W2VLayerBuf& b = m->LB[l];
    void* h_out = (l + 1 < c.layers) ? m->LB[l + 1].h_in : m->enc_out;
    TS_TRY(layernorm_fwd(ctx, dt, b.h_in, nullptr, m->P + o.ln1_g, m->P + o.ln1_b, b.x1, nullptr, b.ln1_mean, b.ln1_rstd, M, H, c.ln_eps, st));
    TS_TRY(GemmB(dt, dt).A(b.x1, 0, H).B(m->W(o.qkv_w), 1, 3 * H).C(b.qkv, 3 * H).bias(m->P + o.qkv_b).mnk(M, 3 * H, H).run(ctx, st));
    const char* qkv = (const char*)b.qkv;
    const float adrop = m->drop(c.attention_dropout);
    if (m->fused_attn) {
      ts_attn_desc a;
      memset(&a, 0, sizeof(a));
      a.q = qkv; a.k = qkv + m->E(H); a.v = qkv + m->E(2 * H); a.o = b.ctx;
      a.q_ld = a.kv_ld = 3 * H; a.q_bs = a.kv_bs = (long long)T * 3 * H; a.o_ld = H; a.o_bs = (long long)T * H;
      a.stats = (float*)b.P; a.batch = B; a.heads = nh; a.tq = a.tk = T; a.head_dim = hd;
      a.scale = 1.f / sqrtf((float)hd); a.mask_mode = 0; a.drop = adrop; a.seed = site_seed(seed, 100 + l * 8); a.o_lo = b.ctx_lo;
      TS_TRY(attn_fwd(ctx, &a, st));
    } else {
      // scores = q k^T (scaled by 1/sqrt(hd) inside the softmax, V:349)
      TS_TRY(GemmB(dt, dt).A(qkv, 0, 3 * H).astride(hd, (long long)T * 3 * H)
                 .B(qkv + m->E(H), 0, 3 * H).bstride(hd, (long long)T * 3 * H)
                 .C(b.P, Tp).cstride((long long)T * Tp, (long long)nh * T * Tp).mnk(T, T, hd).batch(nh, B).run(ctx, st));
      TS_TRY(softmax_fwd(ctx, dt, b.P, Tp, B * nh, T, T, 1.f / sqrtf((float)hd), 0, adrop, site_seed(seed, 100 + l * 8), m->g_Pd, st));
      const void* Puse = adrop > 0 ? m->g_Pd : b.P;
      TS_TRY(GemmB(dt, dt).A(Puse, 0, Tp).astride((long long)T * Tp, (long long)nh * T * Tp)
                 .B(qkv + m->E(2 * H), 1, 3 * H).bstride(hd, (long long)T * 3 * H)
                 .C(b.ctx, H).cstride(hd, (long long)T * H).mnk(T, hd, T).batch(nh, B).run(ctx, st));
    }
    TS_TRY(GemmB(dt, dt).A(b.ctx, 0, H).B(m->W(o.o_w), 1, H).C(b.h_mid, H).bias(m->P + o.o_b).res(b.h_in, H)
               .drop(m->drop(c.hidden_dropout), site_seed(seed, 101 + l * 8)).mnk(M, H, H).run(ctx, st));
    TS_TRY(layernorm_fwd(ctx, dt, b.h_mid, nullptr, m->P + o.ln2_g, m->P + o.ln2_b, b.x2, nullptr, b.ln2_mean, b.ln2_rstd, M, H, c.ln_eps, st));
    TS_TRY(GemmB(dt, dt).A(b.x2, 0, H).B(m->W(o.fc1_w), 1, F).C(b.f, F).bias(m->P + o.fc1_b).gelu(b.u)
               .drop(m->drop(c.activation_dropout), site_seed(seed, 102 + l * 8)).mnk(M, F, H).run(ctx, st));
    TS_TRY(GemmB(dt, dt).A(b.f, 0, F).B(m->W(o.fc2_w), 1, H).C(h_out, H).bias(m->P + o.fc2_b).res(b.h_mid, H)
               .drop(m->drop(c.hidden_dropout), site_seed(seed, 103 + l * 8)).mnk(M, H, F).run(ctx, st));
  }
  if (c.head == 1) {
    // ---- Wav2Vec2ForCTC (V:971-1000): dropout -> lm_head -> mean CE against class 0 on every frame ---------------
    const int V = c.vocab_size;
    const void* hin = m->enc_out;
    if (m->drop(c.hidden_dropout) > 0) { TS_TRY(dropout_apply(ctx, dt, m->enc_out, m->hd_in, (long long)M * H, c.hidden_dropout, site_seed(seed, 5), st)); hin = m->hd_in; }
    TS_TRY(GemmB(dt, TS_F32).A(hin, 0, H).B(m->W(m->lm_w), 1, V).C(m->hlogits, V).bias(m->P + m->lm_b).mnk(M, V, H).simt().run(ctx, st));
    if (with_loss) {
      TS_TRY(ce_rows_fwd_bwd(ctx, dt, m->hlogits, V, nullptr, m->d_hlogits, V, m->scalars + 3, M, V, 1.f / (m->loss_div * (float)M), st));
      ts::launch_k(w2v_finalize_scalars, 1, 1, 0, st, m->scalars, 1.f / (float)M, 0.f);
      TS_LAUNCH_OK(ctx);
    }
    m->fwd_done = with_loss;
    return 0;
  }
  if (c.head == 2) {
    // ---- Wav2Vec2ForSequenceClassification (V:1033-1056): mean over time -> Dense+tanh -> dropout -> Dense -> CE -----
    const int Pc = c.classifier_proj, NL = c.num_labels;
    TS_TRY(mean_pool_fwd(ctx, dt, m->enc_out, m->pooled, B, T, H, st));
    TS_TRY(GemmB(dt, dt).A(m->pooled, 0, H).B(m->W(m->cp_w), 1, Pc).C(m->cp_pre, Pc).bias(m->P + m->cp_b).mnk(B, Pc, H).simt().run(ctx, st));
    TS_TRY(tanh_drop_fwd(ctx, dt, m->cp_pre, m->cp_act, m->cp_drop, (long long)B * Pc, m->drop(c.hidden_dropout), site_seed(seed, 6), st));
    TS_TRY(GemmB(dt, TS_F32).A(m->cp_drop, 0, Pc).B(m->W(m->cl_w), 1, NL).C(m->hlogits, NL).bias(m->P + m->cl_b).mnk(B, NL, Pc).simt().run(ctx, st));
    if (with_loss) {
      TS_TRY(ce_rows_fwd_bwd(ctx, dt, m->hlogits, NL, labels, m->d_hlogits, NL, m->scalars + 3, B, NL, 1.f / (m->loss_div * (float)B), st));
      ts::launch_k(w2v_finalize_scalars, 1, 1, 0, st, m->scalars, 1.f / (float)B, 0.f);
      TS_LAUNCH_OK(ctx);
    }
    m->fwd_done = with_loss;
    return 0;
  }
  // ---- projection heads (V:854-857) ------------------------------------------------------------------------
  TS_TRY(GemmB(dt, dt).A(m->enc_out, 0, H).B(m->W(m->ph_w), 1, Pj).C(m->ph_lin, Pj).bias(m->P + m->ph_b).mnk(M, Pj, H).run(ctx, st));
  TS_TRY(layernorm_fwd(ctx, dt, m->ph_lin, nullptr, m->P + m->ph_ln_g, m->P + m->ph_ln_b, m->ps, nullptr, m->ph_mean, m->ph_rstd, M, Pj, c.ln_eps, st));
  TS_TRY(GemmB(dt, dt).A(m->qfeat, 0, D).B(m->W(m->pq_w), 1, Pj).C(m->pq_lin, Pj).bias(m->P + m->pq_b).mnk(M, Pj, D).run(ctx, st));
  TS_TRY(layernorm_fwd(ctx, dt, m->pq_lin, nullptr, m->P + m->pq_ln_g, m->P + m->pq_ln_b, m->pq, nullptr, m->pq_mean, m->pq_rstd, M, Pj, c.ln_eps, st));
  if (m->drop(c.hidden_dropout) > 0) {
    TS_TRY(dropout_apply(ctx, dt, m->ps, m->ps, (long long)M * Pj, c.hidden_dropout, site_seed(seed, 3), st));
    TS_TRY(dropout_apply(ctx, dt, m->pq, m->pq, (long long)M * Pj, c.hidden_dropout, site_seed(seed, 4), st));
  }
  // ---- contrastive loss over all-pairs similarities (V:865-899) + diversity (V:901-905, V:1220) ---------
  TS_TRY(GemmB(dt, TS_F32).A(m->ps, 0, Pj).astride((long long)T * Pj, 0).B(m->pq, 0, Pj).bstride((long long)T * Pj, 0)
             .C(m->S, Tp).cstride((long long)T * Tp, 0).mnk(T, T, Pj).batch(B, 1).run(ctx, st));
  TS_TRY(contrastive_fwd_bwd(ctx, dt, m->S, Tp, neg, neg_bs, neg_ts, m->dS, Tp, m->logits, m->scalars + 3, B, T,
                             c.num_negatives, c.temperature, 1.f / (m->loss_div * (float)M), st));
  ts::launch_k(w2v_finalize_scalars, 1, 1, 0, st, m->scalars, 1.f / (float)M, c.diversity_weight);
  TS_LAUNCH_OK(ctx);
  m->fwd_done = true;
  return 0;
}

// dense layer backward: dW += X^T dY (fp32; the gradient arena is zeroed at the start of backward, so split-K partials may be added), db += colsum(dY), dX = dY W^T (optional, + residual)
static int dense_bwd(W2V* m, const void* X, int K, const void* dY, int Nn, long long w_off, long long ldw, long long b_off,
                     void* dX, const void* dres, int rows, cudaStream_t st, const void* gelu_u = nullptr, float drop = 0.f,
                     uint64_t drop_seed = 0, bool simt = false) {
  Ctx* ctx = m->ctx;
  const int dt = m->prec;
  GemmB gw(dt, TS_F32);
  gw.A(X, 1, K).B(dY, 1, Nn).C(m->G + w_off, ldw).mnk(K, Nn, rows).acc();
  if (simt) gw.simt();
  TS_TRY(gw.run(ctx, st));
  if (b_off >= 0) TS_TRY(colsum_acc(ctx, dt, dY, Nn, rows, Nn, m->G + b_off, st));
  if (dX) {
    GemmB g(dt, dt);
    g.A(dY, 0, Nn).B(m->W(w_off), 0, ldw).C(dX, K).mnk(rows, K, Nn);
    if (dres) g.res(dres, K);
    if (gelu_u) g.gelu_grad(gelu_u, K).drop(drop, drop_seed);   // dX feeds a GELU (+dropout): its backward runs in the epilogue
    if (simt) g.simt();
    TS_TRY(g.run(ctx, st));
  }
  return 0;
}

static int w2v_backward_stage(W2V* m, int stage, cudaStream_t st) {
  Ctx* ctx = m->ctx;
  const ts_w2v_config& c = m->cfg;
  const int dt = m->prec, B = m->B, n = m->nconv, G = c.pos_groups;
  const int T = m->T, Tp = m->Tp, M = m->M, H = c.hidden, F = c.ffn, nh = c.heads, hd = H / nh;
  const int C = c.conv_dim[n - 1], D = c.cv_dim, Pj = c.proj_dim;
  const uint64_t seed = m->seed;
  const float hdrop = m->drop(c.hidden_dropout);
  if (stage == 0 && c.head == 1) {
    // ---- CTC head: lm_head backward, then back through the dropout on the encoder output (V:975-979) -----------
    const void* hin = hdrop > 0 ? m->hd_in : m->enc_out;
    TS_TRY(dense_bwd(m, hin, H, m->d_hlogits, c.vocab_size, m->lm_w, c.vocab_size, m->lm_b, m->g_a, nullptr, M, st, nullptr, 0.f, 0, true));
    if (hdrop > 0) TS_TRY(dropout_apply(ctx, dt, m->g_a, m->g_a, (long long)M * H, hdrop, site_seed(seed, 5), st));
    return 0;
  }
  if (stage == 0 && c.head == 2) {
    // ---- classification head: classifier -> dropout/tanh -> classifier_proj -> mean pooling (V:1043-1048) ------
    const int Pc = c.classifier_proj, NL = c.num_labels;
    TS_TRY(dense_bwd(m, m->cp_drop, Pc, m->d_hlogits, NL, m->cl_w, NL, m->cl_b, m->d_cp, nullptr, B, st, nullptr, 0.f, 0, true));
    TS_TRY(tanh_drop_bwd(ctx, dt, m->d_cp, m->cp_act, m->d_cp, (long long)B * Pc, hdrop, site_seed(seed, 6), st));
    TS_TRY(dense_bwd(m, m->pooled, H, m->d_cp, Pc, m->cp_w, Pc, m->cp_b, m->d_pooled, nullptr, B, st, nullptr, 0.f, 0, true));
    TS_TRY(mean_pool_bwd(ctx, dt, m->d_pooled, m->g_a, B, T, H, st));   // g_a = d(enc_out)
    return 0;
  }
  if (stage == 0) {
    // ---- heads: contrastive -> projection heads -> codebook --------------------------------------------
    void* dps = m->g_small1;
    void* dpq = m->g_small2;
    TS_TRY(GemmB(dt, dt).A(m->dS, 0, Tp).astride((long long)T * Tp, 0).B(m->pq, 1, Pj).bstride((long long)T * Pj, 0)
               .C(dps, Pj).cstride((long long)T * Pj, 0).mnk(T, Pj, T).batch(B, 1).run(ctx, st));
    TS_TRY(GemmB(dt, dt).A(m->dS, 1, Tp).astride((long long)T * Tp, 0).B(m->ps, 1, Pj).bstride((long long)T * Pj, 0)
               .C(dpq, Pj).cstride((long long)T * Pj, 0).mnk(T, Pj, T).batch(B, 1).run(ctx, st));
    if (hdrop > 0) {
      TS_TRY(dropout_apply(ctx, dt, dps, dps, (long long)M * Pj, hdrop, site_seed(seed, 3), st));
      TS_TRY(dropout_apply(ctx, dt, dpq, dpq, (long long)M * Pj, hdrop, site_seed(seed, 4), st));
    }
    // project_q: LN bwd -> dense bwd -> d(quantized) -> codebook scatter (no gradient into the VQ input, V:631-638)
    TS_TRY(layernorm_bwd(ctx, dt, dpq, m->pq_lin, m->P + m->pq_ln_g, m->pq_mean, m->pq_rstd, nullptr, dpq, m->G + m->pq_ln_g, m->G + m->pq_ln_b, M, Pj, st));
    void* dq = m->g_x;
    TS_TRY(dense_bwd(m, m->qfeat, D, dpq, Pj, m->pq_w, Pj, m->pq_b, dq, nullptr, M, st));
    TS_TRY(vq_bwd(ctx, dt, dq, m->code_idx, m->G + m->cb, M, c.cv_groups, c.cv_per_group, D / c.cv_groups, st));
    // project_hid
    TS_TRY(layernorm_bwd(ctx, dt, dps, m->ph_lin, m->P + m->ph_ln_g, m->ph_mean, m->ph_rstd, nullptr, dps, m->G + m->ph_ln_g, m->G + m->ph_ln_b, M, Pj, st));
    TS_TRY(dense_bwd(m, m->enc_out, H, dps, Pj, m->ph_w, Pj, m->ph_b, m->g_a, nullptr, M, st));  // g_a = d(enc_out)
    return 0;
  }
  if (stage >= 1 && stage <= c.layers) {
    const int l = c.layers - stage;
    const W2VLayerOff& o = m->L[l];
    W2VLayerBuf& b = m->LB[l];
    void* dh = m->g_a;        // gradient wrt this layer's output; final dh_in is written back to g_a
    // Every bias gradient below is the column sum of a tensor some element-wise pass has just produced, so that pass takes it along
    // (dropout_colsum / gelu_bwd_colsum / layernorm_bwd_drop) and dense_bwd is told not to re-read dY (bias offset -1).
    void* t1 = dh;
    if (hdrop > 0) { t1 = m->g_t; TS_TRY(dropout_colsum(ctx, dt, dh, t1, M, H, m->G + o.fc2_b, hdrop, site_seed(seed, 103 + l * 8), st)); }
    // fc2
    // (the GEMM engine can apply GELU'(u) o dropout in the dgrad epilogue — ts_gemm_desc.act = 2 — but with K = 768 that GEMM is
    // epilogue-bound and the separate HBM-bound pass is faster: measured 534 us vs 375 us per step)
    TS_TRY(dense_bwd(m, b.f, F, t1, H, o.fc2_w, H, hdrop > 0 ? -1 : o.fc2_b, m->g_f, nullptr, M, st));
    TS_TRY(gelu_bwd_colsum(ctx, dt, m->g_f, b.u, m->g_f, M, F, m->G + o.fc1_b, m->drop(c.activation_dropout), site_seed(seed, 102 + l * 8), st));
    // fc1
    TS_TRY(dense_bwd(m, b.x2, H, m->g_f, F, o.fc1_w, F, -1, m->g_x, nullptr, M, st));
    // LN2: dh_mid = dh + LNbwd(dx2); the same pass emits t2 = dropout(dh_mid) (the Dropout behind out_proj, V:431) and its column sums
    void* dh_mid = m->g_b;
    void* t2 = dh_mid;
    if (hdrop > 0) {
      t2 = m->g_t;
      TS_TRY(layernorm_bwd_drop(ctx, dt, m->g_x, b.h_mid, m->P + o.ln2_g, b.ln2_mean, b.ln2_rstd, dh, dh_mid, m->G + o.ln2_g, m->G + o.ln2_b, M, H,
                                t2, m->G + o.o_b, hdrop, site_seed(seed, 101 + l * 8), st));
    } else {
      TS_TRY(layernorm_bwd(ctx, dt, m->g_x, b.h_mid, m->P + o.ln2_g, b.ln2_mean, b.ln2_rstd, dh, dh_mid, m->G + o.ln2_g, m->G + o.ln2_b, M, H, st));
    }
    // out_proj
    TS_TRY(dense_bwd(m, b.ctx, H, t2, H, o.o_w, H, hdrop > 0 ? -1 : o.o_b, m->g_ctx, nullptr, M, st));
    // attention core
    const float adrop = m->drop(c.attention_dropout);
    const uint64_t aseed = site_seed(seed, 100 + l * 8);
    const char* qkv = (const char*)b.qkv;
    char* dqkv = (char*)m->g_qkv;
    if (m->fused_attn) {
      ts_attn_desc a;
      memset(&a, 0, sizeof(a));
      a.q = qkv; a.k = qkv + m->E(H); a.v = qkv + m->E(2 * H); a.o = b.ctx;
      a.q_ld = a.kv_ld = 3 * H; a.q_bs = a.kv_bs = (long long)T * 3 * H; a.o_ld = H; a.o_bs = (long long)T * H;
      a.stats = (float*)b.P; a.batch = B; a.heads = nh; a.tq = a.tk = T; a.head_dim = hd;
      a.scale = 1.f / sqrtf((float)hd); a.mask_mode = 0; a.drop = adrop; a.seed = aseed; a.o_lo = b.ctx_lo;
      a.d_o = m->g_ctx; a.dq = dqkv; a.dk = dqkv + m->E(H); a.dv = dqkv + m->E(2 * H);
      a.dq_ld = a.dkv_ld = 3 * H; a.dq_bs = a.dkv_bs = (long long)T * 3 * H; a.dsum = (float*)m->g_P; a.dq_accum = (float*)m->g_dqacc;
      TS_TRY(attn_bwd(ctx, &a, st));
    } else {
      const void* Puse = b.P;
      if (adrop > 0) { TS_TRY(dropout_apply(ctx, dt, b.P, m->g_Pd, (long long)B * nh * T * Tp, adrop, aseed, st)); Puse = m->g_Pd; }
      const long long sP1 = (long long)T * Tp, sP2 = (long long)nh * T * Tp, sQ2 = (long long)T * 3 * H, sC2 = (long long)T * H;
      // dV = Pd^T dctx
      TS_TRY(GemmB(dt, dt).A(Puse, 1, Tp).astride(sP1, sP2).B(m->g_ctx, 1, H).bstride(hd, sC2)
                 .C(dqkv + m->E(2 * H), 3 * H).cstride(hd, sQ2).mnk(T, hd, T).batch(nh, B).run(ctx, st));
      // dPd = dctx V^T
      TS_TRY(GemmB(dt, dt).A(m->g_ctx, 0, H).astride(hd, sC2).B(qkv + m->E(2 * H), 0, 3 * H).bstride(hd, sQ2)
                 .C(m->g_P, Tp).cstride(sP1, sP2).mnk(T, T, hd).batch(nh, B).run(ctx, st));
      TS_TRY(softmax_bwd(ctx, dt, b.P, m->g_P, Tp, B * nh, T, T, 1.f / sqrtf((float)hd), adrop, aseed, st));
      // dQ = dS K ; dK = dS^T Q
      TS_TRY(GemmB(dt, dt).A(m->g_P, 0, Tp).astride(sP1, sP2).B(qkv + m->E(H), 1, 3 * H).bstride(hd, sQ2)
                 .C(dqkv, 3 * H).cstride(hd, sQ2).mnk(T, hd, T).batch(nh, B).run(ctx, st));
      TS_TRY(GemmB(dt, dt).A(m->g_P, 1, Tp).astride(sP1, sP2).B(qkv, 1, 3 * H).bstride(hd, sQ2)
                 .C(dqkv + m->E(H), 3 * H).cstride(hd, sQ2).mnk(T, hd, T).batch(nh, B).run(ctx, st));
    }
    // qkv projection
    TS_TRY(dense_bwd(m, b.x1, H, dqkv, 3 * H, o.qkv_w, 3 * H, o.qkv_b, m->g_x, nullptr, M, st));
    // LN1: dh_in = dh_mid + LNbwd(dx1)
    TS_TRY(layernorm_bwd(ctx, dt, m->g_x, b.h_in, m->P + o.ln1_g, b.ln1_mean, b.ln1_rstd, dh_mid, m->g_a, m->G + o.ln1_g, m->G + o.ln1_b, M, H, st));
    return 0;
  }
  // ---- front end: feature projection, fe LayerNorm, positional conv, conv stack ----------------------------
  void* dhs = m->g_a;
  if (hdrop > 0) TS_TRY(dropout_apply(ctx, dt, dhs, dhs, (long long)M * H, hdrop, site_seed(seed, 2), st));
  TS_TRY(layernorm_bwd(ctx, dt, dhs, m->fp_out, m->P + m->fp_ln_g, m->fp_mean, m->fp_rstd, nullptr, m->g_b, m->G + m->fp_ln_g, m->G + m->fp_ln_b, M, H, st));
  void* def = m->g_x;
  TS_TRY(dense_bwd(m, m->ef, C, m->g_b, H, m->fp_w, H, m->fp_b, def, nullptr, M, st));
  if (hdrop > 0) TS_TRY(dropout_apply(ctx, dt, def, def, (long long)M * C, hdrop, site_seed(seed, 1), st));
  void* dsum = m->g_t;
  TS_TRY(layernorm_bwd(ctx, dt, def, m->possum, m->P + m->fe_ln_g, m->fe_mean, m->fe_rstd, nullptr, dsum, m->G + m->fe_ln_g, m->G + m->fe_ln_b, M, C, st));
  void* da_last = m->g_b;
  {
    const int K = c.pos_kernel, cpg = C / G, Rp = T + K - 1, left = (K - 1) / 2, leftp = K - 1 - left;
    TS_TRY(colsum_acc(ctx, dt, dsum, C, M, C, m->G + m->pos_b, st));
    TS_TRY(posconv_pack(ctx, dt, dsum, m->g_dyg, B, T, C, G, K, leftp, st));
    TS_TRY(fill_zero(ctx, (char*)m->g_dyg + m->E((long long)G * B * Rp * cpg), m->E((long long)K * cpg), st));
    // wgrad: dW[(j,c), g*cpg+o] = sum_{b,t} hp[b,t+j,c] * dy[b,t,o]  (uniform rows over all batches; dummy rows hit zeros)
    TS_TRY(GemmB(dt, TS_F32).A(m->hg, 1, cpg).astride((long long)B * Rp * cpg, 0)
               .B((const char*)m->g_dyg + m->E((long long)leftp * cpg), 1, cpg).bstride((long long)B * Rp * cpg, 0)
               .C(m->G + m->pos_w, C).cstride(cpg, 0).mnk(K * cpg, cpg, B * Rp - leftp).batch(G, 1).acc().run(ctx, st));
    // dgrad: da[b,tau,g*cpg+c] = dsum + sum_{jj,o} dyp[b,tau+jj,o] * wt[g][jj][o][c]
    TS_TRY(GemmB(dt, dt).A(m->g_dyg, 0, cpg).astride((long long)B * Rp * cpg, (long long)Rp * cpg)
               .B(m->wt_flip, 1, cpg).bstride((long long)K * cpg * cpg, 0)
               .C(da_last, C).cstride(cpg, (long long)T * C).res(dsum, C, cpg, (long long)T * C)
               .mnk(T, cpg, K * cpg).batch(G, B).run(ctx, st));
  }
  const bool fuse_l0 = !m->no_fused_gn;
  for (int i = n - 1; i >= 0; --i) {
    const int Ci = c.conv_dim[i];
    Col2imSrc col;
    const Col2imSrc* colp = nullptr;
    if (i + 1 < n) {
      col.dcol = m->g_dcol; col.rows_per_batch = m->Rq[i + 1]; col.t_next = m->Tc[i + 1];
      col.k = c.conv_kernel[i + 1]; col.s = c.conv_stride[i + 1]; col.left = m->padl[i + 1];
      colp = &col;
    }
    TS_TRY(groupnorm_gelu_bwd(ctx, dt, (i + 1 < n) ? nullptr : da_last, T, colp, m->c[i], m->rpb_c[i], m->gn_mean[i], m->gn_rstd[i],
                              m->P + m->conv_g[i], m->P + m->conv_b[i], m->g_dc, m->rpb_c[i], m->G + m->conv_g[i], m->G + m->conv_b[i],
                              m->gn_accum, B, m->Tc[i], Ci, G, st, /*skip_pass2=*/fuse_l0 && i == 0));
    if (i == 0) {
      // layer 0: the only consumer of d(conv0 output) is this weight gradient, which forms it on load (conv_fe.cu Conv0GnPass2)
      if (fuse_l0)
        TS_TRY(conv0_wgrad(ctx, dt, m->wave, m->g_dc, m->rpb_c[0], m->G + m->conv_w[0], B, m->N, m->Tc[0], Ci, c.conv_kernel[0],
                           c.conv_stride[0], m->padl[0], st, m->c[0], m->rpb_c[0], m->gn_mean[0], m->gn_rstd[0], m->P + m->conv_g[0],
                           m->gn_accum, G));
      else
      TS_TRY(conv0_wgrad(ctx, dt, m->wave, m->g_dc, m->rpb_c[0], m->G + m->conv_w[0], B, m->N, m->Tc[0], Ci, c.conv_kernel[0],
                         c.conv_stride[0], m->padl[0], st));
    } else {
      const int Cin = c.conv_dim[i - 1], k = c.conv_kernel[i], s = c.conv_stride[i];
      const int rows = B * m->Rq[i];
      TS_TRY(GemmB(dt, TS_F32).A(m->a[i - 1], 1, (long long)s * Cin).B(m->g_dc, 1, Ci).C(m->G + m->conv_w[i], Ci)
                 .mnk(k * Cin, Ci, rows).acc().run(ctx, st));
      TS_TRY(GemmB(dt, dt).A(m->g_dc, 0, Ci).B(m->W(m->conv_w[i]), 0, Ci).C(m->g_dcol, (long long)k * Cin)
                 .mnk(rows, k * Cin, Ci).run(ctx, st));
    }
  }
  return 0;
}

}  // namespace ts

using namespace ts;

extern "C" {

int ts_w2v_create(ts_ctx* ctx_, const ts_w2v_config* cfg, int precision, ts_w2v** out) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !cfg || !out) return TS_EINVAL;
  TS_REQUIRE(ctx, precision == TS_F32 || precision == TS_BF16, TS_EDTYPE, "w2v: precision must be TS_F32 or TS_BF16");
  TS_REQUIRE(ctx, cfg->n_conv >= 2 && cfg->n_conv <= 8 && cfg->layers <= 64, TS_EINVAL, "w2v: bad config");
  TS_REQUIRE(ctx, cfg->head >= 0 && cfg->head <= 2, TS_EINVAL, "w2v: head must be 0 (pre-training), 1 (CTC) or 2 (classification)");
  TS_REQUIRE(ctx, cfg->head != 1 || cfg->vocab_size > 0, TS_EINVAL, "w2v: CTC head needs vocab_size > 0");
  TS_REQUIRE(ctx, cfg->head != 2 || (cfg->classifier_proj > 0 && cfg->num_labels > 0), TS_EINVAL,
             "w2v: classification head needs classifier_proj > 0 and num_labels > 0");
  W2V* m = new W2V();
  m->ctx = ctx; m->cfg = *cfg; m->prec = precision; m->esz = precision == TS_BF16 ? 2 : 4;
  m->fused_attn = precision == TS_BF16 && cfg->heads > 0 && cfg->hidden / cfg->heads == 64 && !getenv("TETHYS_UNFUSED_ATTENTION");
  build_params(m);
  *out = reinterpret_cast<ts_w2v*>(m);
  return 0;
}
void ts_w2v_destroy(ts_w2v* h) { delete reinterpret_cast<W2V*>(h); }
int64_t ts_w2v_arena_elems(ts_w2v* h) { return reinterpret_cast<W2V*>(h)->pt.n; }
int ts_w2v_num_params(ts_w2v* h) { return (int)reinterpret_cast<W2V*>(h)->pt.defs.size(); }
int ts_w2v_param_info(ts_w2v* h, int i, char* name, int cap, int64_t* offset, int32_t* ndim, int64_t* shape4, int64_t* ld) {
  W2V* m = reinterpret_cast<W2V*>(h);
  if (i < 0 || i >= (int)m->pt.defs.size()) return TS_EINVAL;
  const ParamDef& d = m->pt.defs[i];
  if (name && cap > 0) { strncpy(name, d.name.c_str(), cap - 1); name[cap - 1] = 0; }
  if (offset) *offset = d.offset;
  if (ndim) *ndim = d.ndim;
  if (shape4) for (int j = 0; j < 4; ++j) shape4[j] = d.shape[j];
  if (ld) *ld = d.ld;
  return 0;
}
int ts_w2v_num_stages(ts_w2v* h) { return (int)reinterpret_cast<W2V*>(h)->stage_end.size(); }
int64_t ts_w2v_stage_end(ts_w2v* h, int stage) {
  W2V* m = reinterpret_cast<W2V*>(h);
  if (stage < 0 || stage >= (int)m->stage_end.size()) return -1;
  return m->stage_end[stage];
}
int64_t ts_w2v_workspace_bytes(ts_w2v* h, int B, int N) {
  W2V* m = reinterpret_cast<W2V*>(h);
  W2V tmp = *m;
  Bump bp;
  if (plan(&tmp, B, N, bp)) { m->ctx->err = tmp.ctx->err; return -1; }
  return (int64_t)bp.off + 4096;
}
int ts_w2v_bind(ts_w2v* h, float* params, float* grads, void* params_lp, void* ws, int64_t ws_bytes) {
  W2V* m = reinterpret_cast<W2V*>(h);
  TS_REQUIRE(m->ctx, params && grads && ws, TS_EINVAL, "w2v_bind: null arena");
  TS_REQUIRE(m->ctx, m->prec != TS_BF16 || params_lp, TS_EINVAL, "w2v_bind: bf16 mode needs the bf16 parameter arena");
  m->P = params; m->G = grads; m->P16 = params_lp; m->ws = (char*)ws; m->ws_bytes = ws_bytes;
  m->planned = false;
  return 0;
}
int ts_w2v_sync_compute_weights(ts_w2v* h, void* stream) {
  W2V* m = reinterpret_cast<W2V*>(h);
  cudaStream_t st = (cudaStream_t)stream;
  if (m->prec == TS_BF16) TS_TRY(cast_f32_to_bf16(m->ctx, m->P, m->P16, m->pt.n, st));
  return 0;
}
int ts_w2v_forward(ts_w2v* h, const float* wave, int B, int N, const int* neg, int64_t neg_bs, int64_t neg_ts, float loss_div,
                   uint64_t seed, int training, void* stream) {
  W2V* m = reinterpret_cast<W2V*>(h);
  Ctx* ctx = m->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  TS_REQUIRE(ctx, m->P && m->ws, TS_EINVAL, "w2v_forward: call ts_w2v_bind first");
  TS_REQUIRE(ctx, B > 0 && N > 0 && wave && neg, TS_EINVAL, "w2v_forward: bad arguments");
  TS_REQUIRE(ctx, m->cfg.head == 0, TS_EINVAL, "w2v_forward: this program has a task head (use ts_w2v_forward_head)");
  if (!m->planned || m->B != B || m->N != N) {
    Bump bp;
    bp.base = m->ws;
    TS_TRY(plan(m, B, N, bp));
    TS_REQUIRE(ctx, (long long)bp.off <= m->ws_bytes, TS_EINVAL, "w2v_forward: workspace too small (%lld < %lld bytes)",
               (long long)m->ws_bytes, (long long)bp.off);
    m->planned = true;
  }
  m->seed = seed; m->training = training; m->loss_div = loss_div;
  // flipped positional-conv kernel for the dgrad window-GEMM (from the compute-dtype weights)
  TS_TRY(posconv_flip_weight(ctx, m->prec, m->W(m->pos_w), m->wt_flip, m->cfg.pos_kernel, m->cfg.conv_dim[m->cfg.n_conv - 1],
                             m->cfg.pos_groups, st));
  return w2v_forward(m, wave, neg, neg_bs, neg_ts, st);
}
int ts_w2v_forward_head(ts_w2v* h, const float* wave, int B, int N, const int* labels, float loss_div, uint64_t seed, int dropout,
                        int training, void* stream) {
  W2V* m = reinterpret_cast<W2V*>(h);
  Ctx* ctx = m->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  TS_REQUIRE(ctx, m->P && m->ws, TS_EINVAL, "w2v_forward_head: call ts_w2v_bind first");
  TS_REQUIRE(ctx, B > 0 && N > 0 && wave, TS_EINVAL, "w2v_forward_head: bad arguments");
  TS_REQUIRE(ctx, m->cfg.head == 1 || m->cfg.head == 2, TS_EINVAL, "w2v_forward_head: this is the pre-training program (use ts_w2v_forward)");
  if (!m->planned || m->B != B || m->N != N) {
    Bump bp;
    bp.base = m->ws;
    TS_TRY(plan(m, B, N, bp));
    TS_REQUIRE(ctx, (long long)bp.off <= m->ws_bytes, TS_EINVAL, "w2v_forward_head: workspace too small (%lld < %lld bytes)",
               (long long)m->ws_bytes, (long long)bp.off);
    m->planned = true;
  }
  m->seed = seed; m->training = (dropout && training) ? 1 : 0; m->loss_div = loss_div > 0.f ? loss_div : 1.f; m->fwd_done = false;
  if (training)
    TS_TRY(posconv_flip_weight(ctx, m->prec, m->W(m->pos_w), m->wt_flip, m->cfg.pos_kernel, m->cfg.conv_dim[m->cfg.n_conv - 1],
                               m->cfg.pos_groups, st));
  return w2v_forward(m, wave, nullptr, 0, 0, st, false, labels, training != 0);
}
int ts_w2v_forward_features(ts_w2v* h, const float* wave, int B, int N, void* stream) {
  W2V* m = reinterpret_cast<W2V*>(h);
  Ctx* ctx = m->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  TS_REQUIRE(ctx, m->P && m->ws, TS_EINVAL, "w2v_forward_features: call ts_w2v_bind first");
  TS_REQUIRE(ctx, B > 0 && N > 0 && wave, TS_EINVAL, "w2v_forward_features: bad arguments");
  if (!m->planned || m->B != B || m->N != N) {
    Bump bp;
    bp.base = m->ws;
    TS_TRY(plan(m, B, N, bp));
    TS_REQUIRE(ctx, (long long)bp.off <= m->ws_bytes, TS_EINVAL, "w2v_forward_features: workspace too small");
    m->planned = true;
  }
  m->seed = 0; m->training = 0; m->loss_div = 1.f; m->fwd_done = false;
  return w2v_forward(m, wave, nullptr, 0, 0, st, true);
}
int ts_w2v_backward(ts_w2v* h, int stage_from, int stage_to, void* stream) {
  W2V* m = reinterpret_cast<W2V*>(h);
  cudaStream_t st = (cudaStream_t)stream;
  TS_REQUIRE(m->ctx, m->fwd_done, TS_EINVAL, "w2v_backward: no forward state");
  const int ns = (int)m->stage_end.size();
  if (stage_from <= 0) TS_TRY(fill_zero(m->ctx, m->G, 4ll * m->pt.n, st));
  for (int s = std::max(0, stage_from); s <= std::min(ns - 1, stage_to); ++s) TS_TRY(w2v_backward_stage(m, s, st));
  return 0;
}
int ts_w2v_step(ts_w2v* h, const float* wave, int B, int N, const int* neg, int64_t neg_bs, int64_t neg_ts, const int* labels,
                const ts_step_args* a, void* stream) {
  W2V* m = reinterpret_cast<W2V*>(h);
  if (!m) return TS_EINVAL;
  Ctx* ctx = m->ctx;
  TS_REQUIRE(ctx, a, TS_EINVAL, "w2v_step: args are required");
  int nranks = 1;
  if (a->comm) {
    int rank = 0, ver = 0, reg = 0;
    TS_TRY(ts_comm_info(a->comm, &nranks, &rank, &ver, &reg));
  }
  // V:1231: the loss (and with it every gradient) is divided by the number of replicas before backward
  if (m->cfg.head == 0)
    TS_TRY(ts_w2v_forward(h, wave, B, N, neg, neg_bs, neg_ts, (float)nranks, a->seed, a->dropout ? 1 : 0, stream));
  else
    TS_TRY(ts_w2v_forward_head(h, wave, B, N, labels, (float)nranks, a->seed, a->dropout ? 1 : 0, 1, stream));
  TS_TRY(ts_w2v_backward(h, 0, 1 << 20, stream));
  return step_reduce_update(ctx, m->P, m->G, m->P16, m->pt.n, m->scalars, /*loss_mean_over_replicas=*/true, a, (cudaStream_t)stream);
}
int ts_w2v_get_buffer(ts_w2v* h, const char* name, void** ptr, int32_t* dtype, int32_t* ndim, int64_t* shape4) {
  W2V* m = reinterpret_cast<W2V*>(h);
  if (!m->planned) return set_err(m->ctx, TS_EINVAL, "w2v_get_buffer: no plan yet");
  const ts_w2v_config& c = m->cfg;
  const std::string s(name);
  const long long B = m->B, T = m->T, C = c.conv_dim[c.n_conv - 1];
  auto set = [&](void* p, int dt, int nd, long long a, long long b, long long d3, long long d4) {
    *ptr = p; *dtype = dt; *ndim = nd; shape4[0] = a; shape4[1] = b; shape4[2] = d3; shape4[3] = d4; return 0; };
  if (s == "scalars") return set(m->scalars, TS_F32, 1, 4, 1, 1, 1);
  if (s == "extract_features") return set(m->ef, m->prec, 3, B, T, C, 1);
  if (s == "hidden_states_in") return set(m->hs, m->prec, 3, B, T, c.hidden, 1);
  if (s == "last_hidden_state") return set(m->enc_out, m->prec, 3, B, T, c.hidden, 1);
  if (s == "quantized_features") return set(m->qfeat, m->prec, 3, B, T, c.cv_dim, 1);
  if (s == "quantizer_input") return set(m->z, m->prec, 3, B, T, c.cv_dim, 1);
  if (s == "code_indices") return set(m->code_idx, TS_I64, 3, c.cv_groups, B, T, 1);
  if (s == "head_logits" && c.head == 1) return set(m->hlogits, TS_F32, 3, B, T, c.vocab_size, 1);
  if (s == "d_head_logits" && c.head == 1) return set(m->d_hlogits, m->prec, 3, B, T, c.vocab_size, 1);   // loss gradient the backward reads
  if (s == "head_logits" && c.head == 2) return set(m->hlogits, TS_F32, 2, B, c.num_labels, 1, 1);
  if (s == "pooled_output" && c.head == 2) return set(m->pooled, m->prec, 2, B, c.hidden, 1, 1);
  if (c.head != 0 && (s == "projected_states" || s == "projected_quantized_features" || s == "contrastive_logits"))
    return set_err(m->ctx, TS_EINVAL, "w2v_get_buffer: '%s' exists in the pre-training program only", name);
  if (s == "projected_states") return set(m->ps, m->prec, 3, B, T, c.proj_dim, 1);
  if (s == "projected_quantized_features") return set(m->pq, m->prec, 3, B, T, c.proj_dim, 1);
  if (s == "contrastive_logits") return set(m->logits, TS_F32, 3, B, T, c.num_negatives + 1, 1);
  if (s == "conv_last") return set(m->a[c.n_conv - 1], m->prec, 3, B, T, C, 1);
  if (s == "conv0_raw") return set(m->c[0], m->prec, 3, B, m->rpb_c[0], c.conv_dim[0], 1);
  return set_err(m->ctx, TS_EINVAL, "w2v_get_buffer: unknown buffer '%s'", name);
}

}  // extern "C"
