// Whisper encoder-decoder train step as a native program over pre-bound arenas: forward of
// WhisperForConditionalGeneration.call(labels=…, training=True) (W:547-616: W:324-372 encoder, W:394-466 decoder,
// W:106-176 attention, W:200-206 FFN), the shifted cross-entropy loss (W:585-600) and the hand-derived backward.
// Bug-compatible with the reference: anti-causal decoder mask with fp32 -1e9 absorption, double label
// shift, untied lm_head, query scaling folded into the score scale (mathematically (Wq x + b) * hd^-0.5).
#include <math.h>
#include <stdlib.h>
#include "program.cuh"

namespace ts {

struct AttnOff { long long qkv_w, qkv_b, o_w, o_b; };                         // self-attention: fused [d,3d] (q,k,v)
struct CrossOff { long long q_w, q_b, kv_w, kv_b, o_w, o_b; };                // cross-attention: q [d,d], fused kv [d,2d]
struct EncLayerOff { AttnOff sa; long long ln1_g, ln1_b, fc1_w, fc1_b, fc2_w, fc2_b, ln2_g, ln2_b; };
struct DecLayerOff { AttnOff sa; CrossOff ca; long long ln1_g, ln1_b, ln2_g, ln2_b, ln3_g, ln3_b, fc1_w, fc1_b, fc2_w, fc2_b; };
struct EncLayerBuf { void *h_in, *x1, *qkv, *P, *ctx, *h_mid, *x2, *u, *f; float *m1, *r1, *m2, *r2; };
struct DecLayerBuf {
  void *g_in, *x1, *qkv, *P, *ctx, *g1, *x2, *q, *kv, *Pc, *ctxc, *g2, *x3, *u, *f;
  float *m1, *r1, *m2, *r2, *m3, *r3;
};

struct Whisper {
  Ctx* ctx = nullptr;
  ts_whisper_config cfg;
  int prec = TS_F32, esz = 4;
  bool fused_attn = false;  // bf16 + head_dim 64: tcgen05 flash kernels instead of GEMM/softmax/GEMM
  ParamTable pt;
  long long conv1_w, conv1_b, conv2_w, conv2_b, enc_ln_g, enc_ln_b, emb, dec_ln_g, dec_ln_b, lm_w;
  long long Vp = 0;  // padded lm_head row stride
  std::vector<EncLayerOff> EL;
  std::vector<DecLayerOff> DL;
  std::vector<long long> stage_end;
  float *P = nullptr, *G = nullptr;
  void* P16 = nullptr;
  char* ws = nullptr;
  long long ws_bytes = 0;
  float* pe_enc = nullptr;  // library-owned constant tables [n_ctx,d], [max_target,d]
  float* pe_dec = nullptr;
  // plan
  int B = 0, Tm = 0, T = 0, Tp = 0, S = 0, Sp = 0;
  long long R0 = 0, Rq1 = 0, Rq2 = 0;
  void *xT, *u1, *a1, *u2, *a2, *enc_out, *dec_out, *logits, *dlogits;
  float *enc_m, *enc_r, *dec_m, *dec_r;
  std::vector<EncLayerBuf> EB;
  std::vector<DecLayerBuf> DB;
  void* h_final;  // encoder residual stream after the last layer (input of the final LN)
  void* g_final;  // decoder residual stream after the last layer
  float* scalars;  // [0] loss, [3] raw sum
  const int* labels = nullptr;
  // scratch
  // side stream for the decoder's weight gradients (dense_bwd): fork / join events, all inside the caller's stream order
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // second branch: the cross-attention K/V projections of the encoder output and their backward — full-size GEMMs that depend only
  // on enc_out / feed only d(enc_out), next to the decoder's chain of few-hundred-row kernels
  cudaStream_t side2 = nullptr;
  cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
  float* s_x32;   // fp32 [B*S, d]: split-K target of the lm_head input gradient (K = vocabulary)
  void *s_a, *s_b, *s_t, *s_x, *s_f, *s_ctx, *s_qkv, *s_P, *s_Pd, *s_dqacc, *s_denc, *s_dq, *s_dkv, *s_dcol, *s_du;
  uint64_t seed = 0;
  int training = 1;
  bool planned = false, fwd_done = false;
  bool gen_ready = false;   // ts_whisper_encode has run on the current plan: encoder output + cross K/V are valid

  const void* W(long long off) const { return prec == TS_BF16 ? (const void*)((const bf16*)P16 + off) : (const void*)(P + off); }
  size_t E(long long n) const { return (size_t)n * esz; }
  float drop(float r) const { return training ? r : 0.f; }
};

static void add_attn(ParamTable& pt, const std::string& p, int d, AttnOff& o) {
  o.o_w = pt.add(p + "out_proj.kernel", {d, d});
  o.o_b = pt.add(p + "out_proj.bias", {d});
  o.qkv_w = pt.add_fused({p + "q_proj.kernel", p + "k_proj.kernel", p + "v_proj.kernel"}, d, d, 3 * d);
  o.qkv_b = pt.add_fused({p + "q_proj.bias", p + "k_proj.bias", p + "v_proj.bias"}, 1, d, 3 * d);
}

static void build_params(Whisper* m) {
  const ts_whisper_config& c = m->cfg;
  ParamTable& pt = m->pt;
  const int d = c.d_model, F = c.d_ff;
  m->Vp = (c.vocab + 7) & ~7;
  // backward-completion order (bucketed all-reduce): lm_head, decoder (top->bottom), embedding, encoder, conv stem
  m->lm_w = pt.add_fused({"lm_head.kernel"}, d, c.vocab, m->Vp);
  m->dec_ln_g = pt.add("decoder.layer_norm.gamma", {d});
  m->dec_ln_b = pt.add("decoder.layer_norm.beta", {d});
  m->stage_end.push_back(pt.n);
  m->DL.resize(c.dec_layers);
  for (int l = c.dec_layers - 1; l >= 0; --l) {
    const std::string p = "decoder.layers." + std::to_string(l) + ".";
    DecLayerOff& o = m->DL[l];
    o.fc2_w = pt.add(p + "feed_forward.fc2.kernel", {F, d});
    o.fc2_b = pt.add(p + "feed_forward.fc2.bias", {d});
    o.fc1_w = pt.add(p + "feed_forward.fc1.kernel", {d, F});
    o.fc1_b = pt.add(p + "feed_forward.fc1.bias", {F});
    o.ln3_g = pt.add(p + "final_layer_norm.gamma", {d});
    o.ln3_b = pt.add(p + "final_layer_norm.beta", {d});
    o.ca.o_w = pt.add(p + "encoder_attn.out_proj.kernel", {d, d});
    o.ca.o_b = pt.add(p + "encoder_attn.out_proj.bias", {d});
    o.ca.q_w = pt.add(p + "encoder_attn.q_proj.kernel", {d, d});
    o.ca.q_b = pt.add(p + "encoder_attn.q_proj.bias", {d});
    o.ca.kv_w = pt.add_fused({p + "encoder_attn.k_proj.kernel", p + "encoder_attn.v_proj.kernel"}, d, d, 2 * d);
    o.ca.kv_b = pt.add_fused({p + "encoder_attn.k_proj.bias", p + "encoder_attn.v_proj.bias"}, 1, d, 2 * d);
    o.ln2_g = pt.add(p + "encoder_attn_layer_norm.gamma", {d});
    o.ln2_b = pt.add(p + "encoder_attn_layer_norm.beta", {d});
    add_attn(pt, p + "self_attn.", d, o.sa);
    o.ln1_g = pt.add(p + "self_attn_layer_norm.gamma", {d});
    o.ln1_b = pt.add(p + "self_attn_layer_norm.beta", {d});
    m->stage_end.push_back(pt.n);
  }
  m->emb = pt.add("decoder.embed_tokens.embeddings", {c.vocab, d});
  m->enc_ln_g = pt.add("encoder.layer_norm.gamma", {d});
  m->enc_ln_b = pt.add("encoder.layer_norm.beta", {d});
  m->stage_end.push_back(pt.n);
  m->EL.resize(c.enc_layers);
  for (int l = c.enc_layers - 1; l >= 0; --l) {
    const std::string p = "encoder.layers." + std::to_string(l) + ".";
    EncLayerOff& o = m->EL[l];
    o.fc2_w = pt.add(p + "feed_forward.fc2.kernel", {F, d});
    o.fc2_b = pt.add(p + "feed_forward.fc2.bias", {d});
    o.fc1_w = pt.add(p + "feed_forward.fc1.kernel", {d, F});
    o.fc1_b = pt.add(p + "feed_forward.fc1.bias", {F});
    o.ln2_g = pt.add(p + "final_layer_norm.gamma", {d});
    o.ln2_b = pt.add(p + "final_layer_norm.beta", {d});
    add_attn(pt, p + "self_attn.", d, o.sa);
    o.ln1_g = pt.add(p + "self_attn_layer_norm.gamma", {d});
    o.ln1_b = pt.add(p + "self_attn_layer_norm.beta", {d});
    m->stage_end.push_back(pt.n);
  }
  m->conv2_w = pt.add("encoder.conv2.kernel", {3, d, d});
  m->conv2_b = pt.add("encoder.conv2.bias", {d});
  m->conv1_w = pt.add("encoder.conv1.kernel", {3, c.n_mels, d});
  m->conv1_b = pt.add("encoder.conv1.bias", {d});
  pt.n = (pt.n + 63) & ~63ll;
  m->stage_end.push_back(pt.n);
}

// fused attention keeps, per site, the row statistics [B, nh, Tq, 2] fp32 followed by the bf16 rounding residual of the
// context tensor [B, Tq, d] (ts_attn_desc.o_lo)
static inline size_t fused_stats_bytes(int B, int nh, int Tq) { return ((size_t)8 * B * nh * Tq + 255) & ~(size_t)255; }
static inline size_t fused_aux_bytes(int B, int nh, int Tq, int d) { return fused_stats_bytes(B, nh, Tq) + (size_t)2 * B * Tq * d; }

static int plan(Whisper* m, int B, int Tm, int S, Bump& bp) {
  Ctx* ctx = m->ctx;
  const ts_whisper_config& c = m->cfg;
  const int d = c.d_model, F = c.d_ff, nh = c.heads;
  TS_REQUIRE(ctx, Tm % 2 == 0 && Tm >= 4, TS_ESHAPE, "whisper: mel length %d must be even", Tm);
  TS_REQUIRE(ctx, Tm / 2 <= c.n_ctx && S <= c.max_target && S >= 2, TS_ESHAPE, "whisper: T=%d > n_ctx=%d or S=%d > %d", Tm / 2, c.n_ctx, S, c.max_target);
  TS_REQUIRE(ctx, d % nh == 0 && (d / nh) % 8 == 0 && c.n_mels % 8 == 0, TS_ESHAPE, "whisper: d_model/heads/n_mels alignment");
  m->B = B; m->Tm = Tm; m->T = Tm / 2; m->S = S;
  m->Tp = (m->T + 7) & ~7; m->Sp = (S + 7) & ~7;
  m->Rq1 = Tm + 4; m->R0 = m->Rq1; m->Rq2 = m->Rq1 / 2;  // conv1 k3 s1 pads (1,1); conv2 k3 s2 pads (0,1)
  const int T = m->T, Tp = m->Tp, Sp = m->Sp;
  const long long Me = (long long)B * T, Md = (long long)B * S;
  m->xT = bp.get(m->E((long long)B * m->R0 * c.n_mels + 3 * c.n_mels));
  m->u1 = bp.get(m->E((long long)B * m->Rq1 * d));
  m->a1 = bp.get(m->E((long long)B * m->Rq1 * d + 3 * d));
  m->u2 = bp.get(m->E((long long)B * m->Rq2 * d));
  m->a2 = bp.get(m->E((long long)B * m->Rq2 * d));
  m->scalars = (float*)bp.get(64);
  m->EB.resize(c.enc_layers);
  void* h = bp.get(m->E(Me * d));
  for (int l = 0; l < c.enc_layers; ++l) {
    EncLayerBuf& b = m->EB[l];
    b.h_in = h;
    b.x1 = bp.get(m->E(Me * d)); b.qkv = bp.get(m->E(Me * 3 * d));
    b.P = m->fused_attn ? bp.get(fused_aux_bytes(B, nh, T, d)) : bp.get(m->E((long long)B * nh * T * Tp));   // fused: row statistics + O residual
    b.ctx = bp.get(m->E(Me * d)); b.h_mid = bp.get(m->E(Me * d)); b.x2 = bp.get(m->E(Me * d));
    b.u = bp.get(m->E(Me * F)); b.f = bp.get(m->E(Me * F));
    b.m1 = (float*)bp.get(4 * Me); b.r1 = (float*)bp.get(4 * Me); b.m2 = (float*)bp.get(4 * Me); b.r2 = (float*)bp.get(4 * Me);
    h = bp.get(m->E(Me * d));
  }
  m->h_final = h;
  m->enc_out = bp.get(m->E(Me * d));
  m->enc_m = (float*)bp.get(4 * Me); m->enc_r = (float*)bp.get(4 * Me);
  m->DB.resize(c.dec_layers);
  void* g = bp.get(m->E(Md * d));
  for (int l = 0; l < c.dec_layers; ++l) {
    DecLayerBuf& b = m->DB[l];
    b.g_in = g;
    b.x1 = bp.get(m->E(Md * d)); b.qkv = bp.get(m->E(Md * 3 * d)); b.P = m->fused_attn ? bp.get(fused_aux_bytes(B, nh, S, d)) : bp.get(m->E((long long)B * nh * S * Sp));
    b.ctx = bp.get(m->E(Md * d)); b.g1 = bp.get(m->E(Md * d)); b.x2 = bp.get(m->E(Md * d)); b.q = bp.get(m->E(Md * d));
    b.kv = bp.get(m->E(Me * 2 * d)); b.Pc = m->fused_attn ? bp.get(fused_aux_bytes(B, nh, S, d)) : bp.get(m->E((long long)B * nh * S * Tp)); b.ctxc = bp.get(m->E(Md * d));
    b.g2 = bp.get(m->E(Md * d)); b.x3 = bp.get(m->E(Md * d)); b.u = bp.get(m->E(Md * F)); b.f = bp.get(m->E(Md * F));
    b.m1 = (float*)bp.get(4 * Md); b.r1 = (float*)bp.get(4 * Md); b.m2 = (float*)bp.get(4 * Md); b.r2 = (float*)bp.get(4 * Md);
    b.m3 = (float*)bp.get(4 * Md); b.r3 = (float*)bp.get(4 * Md);
    g = bp.get(m->E(Md * d));
  }
  m->g_final = g;
  m->dec_out = bp.get(m->E(Md * d));
  m->dec_m = (float*)bp.get(4 * Md); m->dec_r = (float*)bp.get(4 * Md);
  m->logits = bp.get(m->E(Md * m->Vp));
  m->dlogits = bp.get(m->E(Md * m->Vp));
  // scratch (sized for the encoder, which has the larger row count)
  const long long Mx = std::max(Me, Md);
  m->s_a = bp.get(m->E(Mx * d)); m->s_b = bp.get(m->E(Mx * d)); m->s_t = bp.get(m->E(Mx * d)); m->s_x = bp.get(m->E(Mx * d));
  m->s_x32 = (float*)bp.get(sizeof(float) * (size_t)Md * d);
  m->s_f = bp.get(m->E(Mx * F)); m->s_ctx = bp.get(m->E(Mx * d)); m->s_qkv = bp.get(m->E(Mx * 3 * d));
  const long long pmax = std::max((long long)B * nh * T * Tp, std::max((long long)B * nh * S * Tp, (long long)B * nh * S * Sp));
  if (m->fused_attn) { m->s_P = bp.get(4ll * B * nh * std::max(T, S)); m->s_Pd = nullptr; m->s_dqacc = bp.get(4ll * B * std::max(T, S) * d); }   // fused: D scratch + fp32 dQ accumulator
  else { m->s_P = bp.get(m->E(pmax)); m->s_Pd = bp.get(m->E(pmax)); m->s_dqacc = nullptr; }
  m->s_denc = bp.get(m->E(Me * d)); m->s_dq = bp.get(m->E(Md * d)); m->s_dkv = bp.get(m->E(Me * 2 * d));
  m->s_dcol = bp.get(m->E((long long)B * m->Rq2 * 3 * d));
  m->s_du = bp.get(m->E((long long)B * m->Rq1 * d));
  return 0;
}

struct AttnShape { int B, nh, hd, Tq, Tk, Tkp; };

// softmax(scale * q k^T + mask) v for all (batch, head) pairs; q rows have stride ldq, k/v rows stride ldkv, ctx stride ldc
static int attn_forward(Whisper* m, const void* q, long long ldq, const void* k, const void* v, long long ldkv, void* P, void* ctxo,
                        long long ldc, AttnShape s, float scale, int mask, float drop, uint64_t seed, cudaStream_t st) {
  Ctx* ctx = m->ctx;
  const int dt = m->prec;
  if (m->fused_attn) {  // tcgen05 flash kernel: P is only the [B, nh, Tq, 2] row-statistics buffer
    ts_attn_desc a;
    memset(&a, 0, sizeof(a));
    a.q = q; a.k = k; a.v = v; a.o = ctxo;
    a.q_ld = ldq; a.q_bs = (long long)s.Tq * ldq; a.kv_ld = ldkv; a.kv_bs = (long long)s.Tk * ldkv; a.o_ld = ldc; a.o_bs = (long long)s.Tq * ldc;
    a.stats = (float*)P; a.batch = s.B; a.heads = s.nh; a.tq = s.Tq; a.tk = s.Tk; a.head_dim = s.hd;
    a.scale = scale; a.mask_mode = mask; a.drop = drop; a.seed = seed;
    a.o_lo = (char*)P + fused_stats_bytes(s.B, s.nh, s.Tq);
    return attn_fwd(ctx, &a, st);
  }
  const long long sP1 = (long long)s.Tq * s.Tkp, sP2 = (long long)s.nh * s.Tq * s.Tkp;
  TS_TRY(GemmB(dt, dt).A(q, 0, ldq).astride(s.hd, (long long)s.Tq * ldq).B(k, 0, ldkv).bstride(s.hd, (long long)s.Tk * ldkv)
             .C(P, s.Tkp).cstride(sP1, sP2).mnk(s.Tq, s.Tk, s.hd).batch(s.nh, s.B).run(ctx, st));
  TS_TRY(softmax_fwd(ctx, dt, P, s.Tkp, s.B * s.nh, s.Tq, s.Tk, scale, mask, drop, seed, m->s_Pd, st));
  const void* Puse = drop > 0 ? m->s_Pd : P;
  TS_TRY(GemmB(dt, dt).A(Puse, 0, s.Tkp).astride(sP1, sP2).B(v, 1, ldkv).bstride(s.hd, (long long)s.Tk * ldkv)
             .C(ctxo, ldc).cstride(s.hd, (long long)s.Tq * ldc).mnk(s.Tq, s.hd, s.Tk).batch(s.nh, s.B).run(ctx, st));
  return 0;
}

static int attn_backward(Whisper* m, const void* q, long long ldq, const void* k, const void* v, long long ldkv, const void* P,
                         const void* dctx, long long ldc, void* dq, long long lddq, void* dk, void* dv, long long lddkv, AttnShape s,
                         float scale, float drop, uint64_t seed, cudaStream_t st, const void* ctx_out = nullptr, int mask = 0) {
  Ctx* ctx = m->ctx;
  const int dt = m->prec;
  if (m->fused_attn) {
    ts_attn_desc a;
    memset(&a, 0, sizeof(a));
    a.q = q; a.k = k; a.v = v; a.o = const_cast<void*>(ctx_out);
    a.q_ld = ldq; a.q_bs = (long long)s.Tq * ldq; a.kv_ld = ldkv; a.kv_bs = (long long)s.Tk * ldkv; a.o_ld = ldc; a.o_bs = (long long)s.Tq * ldc;
    a.stats = (float*)const_cast<void*>(P); a.batch = s.B; a.heads = s.nh; a.tq = s.Tq; a.tk = s.Tk; a.head_dim = s.hd;
    a.scale = scale; a.mask_mode = mask; a.drop = drop; a.seed = seed;
    a.d_o = dctx; a.dq = dq; a.dk = dk; a.dv = dv; a.dq_ld = lddq; a.dq_bs = (long long)s.Tq * lddq; a.dkv_ld = lddkv; a.dkv_bs = (long long)s.Tk * lddkv;
    a.dsum = (float*)m->s_P;
    a.dq_accum = (float*)m->s_dqacc;
    a.o_lo = (char*)const_cast<void*>(P) + fused_stats_bytes(s.B, s.nh, s.Tq);
    return attn_bwd(ctx, &a, st);
  }
  const long long sP1 = (long long)s.Tq * s.Tkp, sP2 = (long long)s.nh * s.Tq * s.Tkp;
  const void* Puse = P;
  if (drop > 0) { TS_TRY(dropout_apply(ctx, dt, P, m->s_Pd, (long long)s.B * s.nh * s.Tq * s.Tkp, drop, seed, st)); Puse = m->s_Pd; }
  // dV = Pd^T dctx
  TS_TRY(GemmB(dt, dt).A(Puse, 1, s.Tkp).astride(sP1, sP2).B(dctx, 1, ldc).bstride(s.hd, (long long)s.Tq * ldc)
             .C(dv, lddkv).cstride(s.hd, (long long)s.Tk * lddkv).mnk(s.Tk, s.hd, s.Tq).batch(s.nh, s.B).run(ctx, st));
  // dPd = dctx V^T
  TS_TRY(GemmB(dt, dt).A(dctx, 0, ldc).astride(s.hd, (long long)s.Tq * ldc).B(v, 0, ldkv).bstride(s.hd, (long long)s.Tk * ldkv)
             .C(m->s_P, s.Tkp).cstride(sP1, sP2).mnk(s.Tq, s.Tk, s.hd).batch(s.nh, s.B).run(ctx, st));
  TS_TRY(softmax_bwd(ctx, dt, P, m->s_P, s.Tkp, s.B * s.nh, s.Tq, s.Tk, scale, drop, seed, st));
  // dQ = dS K ; dK = dS^T Q
  TS_TRY(GemmB(dt, dt).A(m->s_P, 0, s.Tkp).astride(sP1, sP2).B(k, 1, ldkv).bstride(s.hd, (long long)s.Tk * ldkv)
             .C(dq, lddq).cstride(s.hd, (long long)s.Tq * lddq).mnk(s.Tq, s.hd, s.Tk).batch(s.nh, s.B).run(ctx, st));
  TS_TRY(GemmB(dt, dt).A(m->s_P, 1, s.Tkp).astride(sP1, sP2).B(q, 1, ldq).bstride(s.hd, (long long)s.Tq * ldq)
             .C(dk, lddkv).cstride(s.hd, (long long)s.Tk * lddkv).mnk(s.Tk, s.hd, s.Tq).batch(s.nh, s.B).run(ctx, st));
  return 0;
}

// dW += X^T dY (fp32; gradient arena zeroed at the start of backward), db += colsum(dY), dX = dY W^T (+ dres)
static int dense_bwd(Whisper* m, const void* X, int K, const void* dY, int Nn, long long w_off, long long ldw, long long b_off,
                     void* dX, const void* dres, long long rows, cudaStream_t st, const void* gelu_u = nullptr, float drop = 0.f,
                     uint64_t drop_seed = 0) {
  Ctx* ctx = m->ctx;
  const int dt = m->prec;
  // Decoder-sized problems (a few hundred rows) occupy a third of the SMs and are mostly launch head and tail: the weight gradient
  // (and the bias column sums) run on a side stream next to the input gradient and are joined right after it — inside a CUDA-graph
  // capture this becomes a fork / join of two branches. Large problems fill the machine on their own and stay in line.
  const bool fork = m->side && dX && rows <= 2048;
  cudaStream_t ws = st;
  if (fork) {
    TS_CUDA_OK(ctx, cudaEventRecord(m->ev_fork, st));
    TS_CUDA_OK(ctx, cudaStreamWaitEvent(m->side, m->ev_fork, 0));
    ws = m->side;
  }
  TS_TRY(GemmB(dt, TS_F32).A(X, 1, K).B(dY, 1, Nn).C(m->G + w_off, ldw).mnk(K, Nn, (int)rows).acc().run(ctx, ws));
  if (b_off >= 0) TS_TRY(colsum_acc(ctx, dt, dY, Nn, (int)rows, Nn, m->G + b_off, ws));
  if (fork) TS_CUDA_OK(ctx, cudaEventRecord(m->ev_join, m->side));
  struct Join {   // the join is queued on every exit path below
    Whisper* m; cudaStream_t st; bool on;
    ~Join() { if (on) cudaStreamWaitEvent(st, m->ev_join, 0); }
  } join{m, st, fork};
  if (dX) {
    GemmB g(dt, dt);
    g.A(dY, 0, Nn).B(m->W(w_off), 0, ldw).C(dX, K).mnk((int)rows, K, Nn);
    if (dres) g.res(dres, K);
    if (gelu_u) g.gelu_grad(gelu_u, K).drop(drop, drop_seed);   // backward of the GELU (+dropout) this gradient feeds, in the epilogue
    TS_TRY(g.run(ctx, st));
  }
  return 0;
}

// [B, n_mels, Tm] fp32 -> rows [B, R0, n_mels] (act dtype), data at rows [1, 1+Tm), the rest zero (W:329 + SAME pad of conv1)
template <typename T>
__global__ void mel_to_rows_kernel(const float* __restrict__ f, T* __restrict__ y, int nm, int Tm, long long R0) {
  ts::pdl_enter();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < nm && t < Tm) ? f[((long long)b * nm + c) * Tm + t] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < Tm && c < nm) y[((long long)b * R0 + 1 + t) * nm + c] = from_f<T>(tile[tx][i]);
  }
}

__global__ void whisper_finalize_scalars(float* s, float inv_rows) {
  ts::pdl_enter(); s[0] = s[3] * inv_rows; }

// conv stem + encoder layers + final LayerNorm -> enc_out (WhisperEncoder.call, W:324-372)
static int whisper_encoder_forward(Whisper* m, const float* feats, cudaStream_t st) {
  Ctx* ctx = m->ctx;
  const ts_whisper_config& c = m->cfg;
  const int dt = m->prec, B = m->B, T = m->T, Tp = m->Tp, d = c.d_model, F = c.d_ff, nh = c.heads, hd = d / nh;
  const long long Me = (long long)B * T;
  const uint64_t seed = m->seed;
  const float scale = 1.f / sqrtf((float)hd);
  TS_TRY(fill_zero(ctx, m->scalars, 64, st));
  // ---- conv stem (W:329-342) ------------------------------------------------------------------------------------
  TS_TRY(fill_zero(ctx, m->xT, m->E((long long)B * m->R0 * c.n_mels + 3 * c.n_mels), st));
  {
    dim3 grid(cdiv(m->Tm, 32), cdiv(c.n_mels, 32), B);
    if (dt == TS_F32) ts::launch_k(mel_to_rows_kernel<float>, grid, 256, 0, st, feats, (float*)m->xT, c.n_mels, m->Tm, m->R0);
    else ts::launch_k(mel_to_rows_kernel<bf16>, grid, 256, 0, st, feats, (bf16*)m->xT, c.n_mels, m->Tm, m->R0);
    TS_LAUNCH_OK(ctx);
  }
  TS_TRY(GemmB(dt, dt).A(m->xT, 0, c.n_mels).B(m->W(m->conv1_w), 1, d).C(m->a1, d).bias(m->P + m->conv1_b).gelu(m->u1)
             .mnk((int)(B * m->Rq1), d, 3 * c.n_mels).run(ctx, st));
  TS_TRY(zero_rows(ctx, dt, m->a1, m->Rq1, m->Tm, B, d, st));
  TS_TRY(fill_zero(ctx, (char*)m->a1 + m->E((long long)B * m->Rq1 * d), m->E(3 * d), st));
  TS_TRY(GemmB(dt, dt).A(m->a1, 0, 2 * d).B(m->W(m->conv2_w), 1, d).C(m->a2, d).bias(m->P + m->conv2_b).gelu(m->u2)
             .mnk((int)(B * m->Rq2), d, 3 * d).run(ctx, st));
  void* h0 = c.enc_layers ? m->EB[0].h_in : m->h_final;
  TS_TRY(add_pe_rows(ctx, dt, m->a2, m->Rq2, m->pe_enc, h0, B, T, d, m->drop(c.dropout), site_seed(seed, 1), st));
  // ---- encoder layers (W:218-236) -------------------------------------------------------------------------------
  for (int l = 0; l < c.enc_layers; ++l) {
    const EncLayerOff& o = m->EL[l];
    EncLayerBuf& b = m->EB[l];
    void* h_out = (l + 1 < c.enc_layers) ? m->EB[l + 1].h_in : m->h_final;
    TS_TRY(layernorm_fwd(ctx, dt, b.h_in, nullptr, m->P + o.ln1_g, m->P + o.ln1_b, b.x1, nullptr, b.m1, b.r1, (int)Me, d, c.ln_eps, st));
    TS_TRY(GemmB(dt, dt).A(b.x1, 0, d).B(m->W(o.sa.qkv_w), 1, 3 * d).C(b.qkv, 3 * d).bias(m->P + o.sa.qkv_b).mnk((int)Me, 3 * d, d).run(ctx, st));
    const char* qkv = (const char*)b.qkv;
    TS_TRY(attn_forward(m, qkv, 3 * d, qkv + m->E(d), qkv + m->E(2 * d), 3 * d, b.P, b.ctx, d, {B, nh, hd, T, T, Tp}, scale, 0,
                        m->drop(c.attention_dropout), site_seed(seed, 100 + l * 8), st));
    TS_TRY(GemmB(dt, dt).A(b.ctx, 0, d).B(m->W(o.sa.o_w), 1, d).C(b.h_mid, d).bias(m->P + o.sa.o_b).res(b.h_in, d).mnk((int)Me, d, d).run(ctx, st));
    TS_TRY(layernorm_fwd(ctx, dt, b.h_mid, nullptr, m->P + o.ln2_g, m->P + o.ln2_b, b.x2, nullptr, b.m2, b.r2, (int)Me, d, c.ln_eps, st));
    TS_TRY(GemmB(dt, dt).A(b.x2, 0, d).B(m->W(o.fc1_w), 1, F).C(b.f, F).bias(m->P + o.fc1_b).gelu(b.u)
               .drop(m->drop(c.activation_dropout), site_seed(seed, 102 + l * 8)).mnk((int)Me, F, d).run(ctx, st));
    TS_TRY(GemmB(dt, dt).A(b.f, 0, F).B(m->W(o.fc2_w), 1, d).C(h_out, d).bias(m->P + o.fc2_b).res(b.h_mid, d)
               .drop(m->drop(c.dropout), site_seed(seed, 103 + l * 8)).mnk((int)Me, d, F).run(ctx, st));
  }
  TS_TRY(layernorm_fwd(ctx, dt, m->h_final, nullptr, m->P + m->enc_ln_g, m->P + m->enc_ln_b, m->enc_out, nullptr, m->enc_m, m->enc_r, (int)Me, d, c.ln_eps, st));
  return 0;
}

// WhisperDecoder.call + lm_head over S positions (S <= the planned length; the buffers are dense [B*S, ...], so a shorter
// sequence uses a prefix of each). Decoder input ids = [start_token, labels[b, 0 .. S-2]] with `label_ld` ints per batch row
// (W:559-563 for the train step; the growing decoder_input_ids of generate(), W:659-704).
//   reuse_cross_kv: the cross-attention K/V projections of the encoder output are already in the layer buffers
//   last_only:      lm_head on the last position only -> logits [B, Vp]
static int whisper_decoder_forward(Whisper* m, const int* labels, long long label_ld, int S, bool reuse_cross_kv, bool last_only,
                                   cudaStream_t st) {
  Ctx* ctx = m->ctx;
  const ts_whisper_config& c = m->cfg;
  const int dt = m->prec, B = m->B, T = m->T, Tp = m->Tp, Sp = (S + 7) & ~7, d = c.d_model, F = c.d_ff, nh = c.heads, hd = d / nh;
  const long long Me = (long long)B * T, Md = (long long)B * S;
  const uint64_t seed = m->seed;
  const float scale = 1.f / sqrtf((float)hd);
  // ---- decoder (W:394-466) --------------------------------------------------------------------------------------
  void* g0 = c.dec_layers ? m->DB[0].g_in : m->g_final;
  // the cross-attention K/V projections depend on the encoder output only: all layers' projections are queued on the second side
  // stream here and joined before the first cross-attention
  const bool kv_side = m->side2 && !reuse_cross_kv && c.dec_layers > 0;
  if (kv_side) {
    TS_CUDA_OK(ctx, cudaEventRecord(m->ev_fork2, st));
    TS_CUDA_OK(ctx, cudaStreamWaitEvent(m->side2, m->ev_fork2, 0));
    for (int l = 0; l < c.dec_layers; ++l) {
      const DecLayerOff& o = m->DL[l];
      TS_TRY(GemmB(dt, dt).A(m->enc_out, 0, d).B(m->W(o.ca.kv_w), 1, 2 * d).C(m->DB[l].kv, 2 * d).bias(m->P + o.ca.kv_b).mnk((int)Me, 2 * d, d).run(ctx, m->side2));
    }
    TS_CUDA_OK(ctx, cudaEventRecord(m->ev_join2, m->side2));
  }
  struct Join2 {   // joined on every exit path (also when a launch below fails)
    Whisper* m; cudaStream_t st; bool pending;
    void now() { if (pending) { cudaStreamWaitEvent(st, m->ev_join2, 0); pending = false; } }
    ~Join2() { now(); }
  } join2{m, st, kv_side};
  TS_TRY(embed_fwd(ctx, dt, m->W(m->emb), labels, label_ld, m->pe_dec, g0, B, S, d, c.start_token, m->drop(c.dropout), site_seed(seed, 2), st));
  for (int l = 0; l < c.dec_layers; ++l) {
    const DecLayerOff& o = m->DL[l];
    DecLayerBuf& b = m->DB[l];
    void* g_out = (l + 1 < c.dec_layers) ? m->DB[l + 1].g_in : m->g_final;
    // self-attention with the reference's anti-causal additive mask (W:414-418, W:150-154)
    TS_TRY(layernorm_fwd(ctx, dt, b.g_in, nullptr, m->P + o.ln1_g, m->P + o.ln1_b, b.x1, nullptr, b.m1, b.r1, (int)Md, d, c.ln_eps, st));
    TS_TRY(GemmB(dt, dt).A(b.x1, 0, d).B(m->W(o.sa.qkv_w), 1, 3 * d).C(b.qkv, 3 * d).bias(m->P + o.sa.qkv_b).mnk((int)Md, 3 * d, d).run(ctx, st));
    const char* qkv = (const char*)b.qkv;
    TS_TRY(attn_forward(m, qkv, 3 * d, qkv + m->E(d), qkv + m->E(2 * d), 3 * d, b.P, b.ctx, d, {B, nh, hd, S, S, Sp}, scale, 1,
                        m->drop(c.attention_dropout), site_seed(seed, 1000 + l * 8), st));
    TS_TRY(GemmB(dt, dt).A(b.ctx, 0, d).B(m->W(o.sa.o_w), 1, d).C(b.g1, d).bias(m->P + o.sa.o_b).res(b.g_in, d).mnk((int)Md, d, d).run(ctx, st));
    // cross-attention over the encoder output (W:278-290)
    TS_TRY(layernorm_fwd(ctx, dt, b.g1, nullptr, m->P + o.ln2_g, m->P + o.ln2_b, b.x2, nullptr, b.m2, b.r2, (int)Md, d, c.ln_eps, st));
    TS_TRY(GemmB(dt, dt).A(b.x2, 0, d).B(m->W(o.ca.q_w), 1, d).C(b.q, d).bias(m->P + o.ca.q_b).mnk((int)Md, d, d).run(ctx, st));
    if (kv_side) join2.now();
    else if (!reuse_cross_kv)
      TS_TRY(GemmB(dt, dt).A(m->enc_out, 0, d).B(m->W(o.ca.kv_w), 1, 2 * d).C(b.kv, 2 * d).bias(m->P + o.ca.kv_b).mnk((int)Me, 2 * d, d).run(ctx, st));
    const char* kv = (const char*)b.kv;
    TS_TRY(attn_forward(m, b.q, d, kv, kv + m->E(d), 2 * d, b.Pc, b.ctxc, d, {B, nh, hd, S, T, Tp}, scale, 0,
                        m->drop(c.attention_dropout), site_seed(seed, 1001 + l * 8), st));
    TS_TRY(GemmB(dt, dt).A(b.ctxc, 0, d).B(m->W(o.ca.o_w), 1, d).C(b.g2, d).bias(m->P + o.ca.o_b).res(b.g1, d).mnk((int)Md, d, d).run(ctx, st));
    // feed-forward
    TS_TRY(layernorm_fwd(ctx, dt, b.g2, nullptr, m->P + o.ln3_g, m->P + o.ln3_b, b.x3, nullptr, b.m3, b.r3, (int)Md, d, c.ln_eps, st));
    TS_TRY(GemmB(dt, dt).A(b.x3, 0, d).B(m->W(o.fc1_w), 1, F).C(b.f, F).bias(m->P + o.fc1_b).gelu(b.u)
               .drop(m->drop(c.activation_dropout), site_seed(seed, 1002 + l * 8)).mnk((int)Md, F, d).run(ctx, st));
    TS_TRY(GemmB(dt, dt).A(b.f, 0, F).B(m->W(o.fc2_w), 1, d).C(g_out, d).bias(m->P + o.fc2_b).res(b.g2, d)
               .drop(m->drop(c.dropout), site_seed(seed, 1003 + l * 8)).mnk((int)Md, d, F).run(ctx, st));
  }
  TS_TRY(layernorm_fwd(ctx, dt, m->g_final, nullptr, m->P + m->dec_ln_g, m->P + m->dec_ln_b, m->dec_out, nullptr, m->dec_m, m->dec_r, (int)Md, d, c.ln_eps, st));
  // ---- lm_head (W:579) --------------------------------------------------------------------------------------------
  if (last_only)
    TS_TRY(GemmB(dt, dt).A((const char*)m->dec_out + m->E((long long)(S - 1) * d), 0, (long long)S * d).B(m->W(m->lm_w), 1, m->Vp)
               .C(m->logits, m->Vp).mnk(B, (int)m->Vp, d).run(ctx, st));
  else
    TS_TRY(GemmB(dt, dt).A(m->dec_out, 0, d).B(m->W(m->lm_w), 1, m->Vp).C(m->logits, m->Vp).mnk((int)Md, (int)m->Vp, d).run(ctx, st));
  return 0;
}

static int whisper_forward(Whisper* m, const float* feats, const int* labels, cudaStream_t st) {
  m->labels = labels;
  TS_TRY(whisper_encoder_forward(m, feats, st));
  return whisper_decoder_forward(m, labels, m->S, m->S, false, false, st);
}

// next token = argmax over the valid vocabulary columns of one logits row per batch item; ties -> lowest index (tf.argmax)
template <typename T>
__global__ void __launch_bounds__(256) argmax_rows_kernel(const T* __restrict__ x, long long ld, int V, int* __restrict__ out,
                                                          long long out_ld) {
  ts::pdl_enter();
  __shared__ float sv[256];
  __shared__ int si[256];
  const T* row = x + (long long)blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < V; j += 256) {
    const float v = to_f<T>(row[j]);
    if (v > best || (v == best && j < bi)) { best = v; bi = j; }
  }
  sv[threadIdx.x] = best; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float ov = sv[threadIdx.x + o];
      const int oi = si[threadIdx.x + o];
      if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[(long long)blockIdx.x * out_ld] = si[0] == 0x7fffffff ? 0 : si[0];
}

static int whisper_loss(Whisper* m, cudaStream_t st) {
  // dlogits = (softmax - onehot) / (B*(S-1)) for the shifted targets; accumulates the loss
  Ctx* ctx = m->ctx;
  TS_TRY(ce_fwd_bwd(ctx, m->prec, m->logits, m->dlogits, m->Vp, m->labels, m->scalars + 3, m->B, m->S, m->cfg.vocab, 1.f, st));
  ts::launch_k(whisper_finalize_scalars, 1, 1, 0, st, m->scalars, 1.f / (float)(m->B * (m->S - 1)));
  TS_LAUNCH_OK(ctx);
  m->fwd_done = true;
  return 0;
}

static int ffn_bwd(Whisper* m, const void* dh, const void* x_in, const void* u, const void* f, long long fc1_w, long long fc1_b,
                   long long fc2_w, long long fc2_b, void* dx_out, long long rows, uint64_t seed_act, uint64_t seed_out, cudaStream_t st) {
  Ctx* ctx = m->ctx;
  const ts_whisper_config& c = m->cfg;
  const int dt = m->prec, d = c.d_model, F = c.d_ff;
  // the two bias gradients are column sums of tensors the element-wise passes produce anyway: they take them along
  const void* t1 = dh;
  const bool dr = m->drop(c.dropout) > 0;
  if (dr) { TS_TRY(dropout_colsum(ctx, dt, dh, m->s_t, (int)rows, d, m->G + fc2_b, c.dropout, seed_out, st)); t1 = m->s_t; }
  TS_TRY(dense_bwd(m, f, F, t1, d, fc2_w, d, dr ? -1 : fc2_b, m->s_f, nullptr, rows, st));
  TS_TRY(gelu_bwd_colsum(ctx, dt, m->s_f, u, m->s_f, (int)rows, F, m->G + fc1_b, m->drop(c.activation_dropout), seed_act, st));
  TS_TRY(dense_bwd(m, x_in, d, m->s_f, F, fc1_w, F, -1, dx_out, nullptr, rows, st));
  return 0;
}

static int whisper_backward_stage(Whisper* m, int stage, cudaStream_t st) {
  Ctx* ctx = m->ctx;
  const ts_whisper_config& c = m->cfg;
  const int dt = m->prec, B = m->B, T = m->T, Tp = m->Tp, S = m->S, Sp = m->Sp, d = c.d_model, nh = c.heads, hd = d / nh;
  const long long Me = (long long)B * T, Md = (long long)B * S;
  const uint64_t seed = m->seed;
  const float scale = 1.f / sqrtf((float)hd);
  const int Ld = c.dec_layers, Le = c.enc_layers;
  if (stage == 0) {
    // lm_head: dW = dec_out^T dlogits ; d(dec_out) = dlogits W^T ; final decoder LN
    TS_TRY(GemmB(dt, TS_F32).A(m->dec_out, 1, d).B(m->dlogits, 1, m->Vp).C(m->G + m->lm_w, m->Vp).mnk(d, (int)m->Vp, (int)Md).acc().run(ctx, st));
    if (dt == TS_BF16) {
      // [B*S, d] = [B*S, vocab] x [vocab, d]: 4 x 12 output tiles with an 811-k-block reduce dim (vocab 51 865). As a bf16-output
      // GEMM that is 48 CTAs walking the whole vocabulary each (228 us); into an fp32 target the engine may split K over all SMs
      // (TMA reduce-add of the partials) and a 1.2 MB cast finishes it.
      TS_TRY(fill_zero(ctx, m->s_x32, (long long)sizeof(float) * Md * d, st));
      TS_TRY(GemmB(dt, TS_F32).A(m->dlogits, 0, m->Vp).B(m->W(m->lm_w), 0, m->Vp).C(m->s_x32, d).mnk((int)Md, d, (int)m->Vp).acc().run(ctx, st));
      TS_TRY(cast_f32_to_bf16(ctx, m->s_x32, m->s_x, Md * d, st));
    } else {
      TS_TRY(GemmB(dt, dt).A(m->dlogits, 0, m->Vp).B(m->W(m->lm_w), 0, m->Vp).C(m->s_x, d).mnk((int)Md, d, (int)m->Vp).run(ctx, st));
    }
    TS_TRY(layernorm_bwd(ctx, dt, m->s_x, m->g_final, m->P + m->dec_ln_g, m->dec_m, m->dec_r, nullptr, m->s_a, m->G + m->dec_ln_g, m->G + m->dec_ln_b, (int)Md, d, st));
    TS_TRY(fill_zero(ctx, m->s_denc, m->E(Me * d), st));
    return 0;
  }
  if (stage >= 1 && stage <= Ld) {
    const int l = Ld - stage;
    const DecLayerOff& o = m->DL[l];
    DecLayerBuf& b = m->DB[l];
    void* dg = m->s_a;  // gradient wrt the layer output; the layer-input gradient is written back to s_a
    TS_TRY(ffn_bwd(m, dg, b.x3, b.u, b.f, o.fc1_w, o.fc1_b, o.fc2_w, o.fc2_b, m->s_x, Md, site_seed(seed, 1002 + l * 8), site_seed(seed, 1003 + l * 8), st));
    void* dg2 = m->s_b;
    TS_TRY(layernorm_bwd(ctx, dt, m->s_x, b.g2, m->P + o.ln3_g, b.m3, b.r3, dg, dg2, m->G + o.ln3_g, m->G + o.ln3_b, (int)Md, d, st));
    // cross-attention
    TS_TRY(dense_bwd(m, b.ctxc, d, dg2, d, o.ca.o_w, d, o.ca.o_b, m->s_ctx, nullptr, Md, st));
    const char* kv = (const char*)b.kv;
    char* dkv = (char*)m->s_dkv;
    TS_TRY(attn_backward(m, b.q, d, kv, kv + m->E(d), 2 * d, b.Pc, m->s_ctx, d, m->s_dq, d, dkv, dkv + m->E(d), 2 * d, {B, nh, hd, S, T, Tp},
                         scale, m->drop(c.attention_dropout), site_seed(seed, 1001 + l * 8), st, b.ctxc, 0));
    TS_TRY(dense_bwd(m, b.x2, d, m->s_dq, d, o.ca.q_w, d, o.ca.q_b, m->s_x, nullptr, Md, st));
    // d_enc += dkv Wkv^T and the K/V weight gradient: three full-size launches that nothing in the rest of this stage depends on
    // (s_denc is consumed after the decoder, s_dkv is rewritten by the NEXT stage's cross-attention) -> second side stream, joined
    // before the stage returns so that the caller's stream order (buckets, next stage) holds
    const bool kvb_side = m->side2 != nullptr;
    cudaStream_t ks = st;
    if (kvb_side) {
      TS_CUDA_OK(ctx, cudaEventRecord(m->ev_fork2, st));
      TS_CUDA_OK(ctx, cudaStreamWaitEvent(m->side2, m->ev_fork2, 0));
      ks = m->side2;
    }
    TS_TRY(dense_bwd(m, m->enc_out, d, dkv, 2 * d, o.ca.kv_w, 2 * d, o.ca.kv_b, m->s_denc, m->s_denc, Me, ks));
    if (kvb_side) TS_CUDA_OK(ctx, cudaEventRecord(m->ev_join2, m->side2));
    struct JoinB {
      Whisper* m; cudaStream_t st; bool on;
      ~JoinB() { if (on) cudaStreamWaitEvent(st, m->ev_join2, 0); }
    } joinb{m, st, kvb_side};
    void* dg1 = m->s_a;
    TS_TRY(layernorm_bwd(ctx, dt, m->s_x, b.g1, m->P + o.ln2_g, b.m2, b.r2, dg2, dg1, m->G + o.ln2_g, m->G + o.ln2_b, (int)Md, d, st));
    // self-attention
    TS_TRY(dense_bwd(m, b.ctx, d, dg1, d, o.sa.o_w, d, o.sa.o_b, m->s_ctx, nullptr, Md, st));
    const char* qkv = (const char*)b.qkv;
    char* dqkv = (char*)m->s_qkv;
    TS_TRY(attn_backward(m, qkv, 3 * d, qkv + m->E(d), qkv + m->E(2 * d), 3 * d, b.P, m->s_ctx, d, dqkv, 3 * d, dqkv + m->E(d), dqkv + m->E(2 * d), 3 * d,
                         {B, nh, hd, S, S, Sp}, scale, m->drop(c.attention_dropout), site_seed(seed, 1000 + l * 8), st, b.ctx, 1));
    TS_TRY(dense_bwd(m, b.x1, d, dqkv, 3 * d, o.sa.qkv_w, 3 * d, o.sa.qkv_b, m->s_x, nullptr, Md, st));
    // in place over dg1: every thread reads its own dres elements before it writes dx
    TS_TRY(layernorm_bwd(ctx, dt, m->s_x, b.g_in, m->P + o.ln1_g, b.m1, b.r1, dg1, m->s_a, m->G + o.ln1_g, m->G + o.ln1_b, (int)Md, d, st));
    return 0;
  }
  if (stage == Ld + 1) {
    // token embedding (scatter-add; dense gradient keeps Adam identical to the IndexedSlices path, App. A-12) and
    // the encoder's final LayerNorm, which receives the accumulated cross-attention gradient
    TS_TRY(embed_bwd(ctx, dt, m->s_a, m->labels, m->G + m->emb, B, S, d, c.start_token, m->drop(c.dropout), site_seed(seed, 2), st));
    TS_TRY(layernorm_bwd(ctx, dt, m->s_denc, m->h_final, m->P + m->enc_ln_g, m->enc_m, m->enc_r, nullptr, m->s_a, m->G + m->enc_ln_g, m->G + m->enc_ln_b, (int)Me, d, st));
    return 0;
  }
  if (stage >= Ld + 2 && stage <= Ld + 1 + Le) {
    const int l = Le - (stage - Ld - 1);
    const EncLayerOff& o = m->EL[l];
    EncLayerBuf& b = m->EB[l];
    void* dh = m->s_a;
    TS_TRY(ffn_bwd(m, dh, b.x2, b.u, b.f, o.fc1_w, o.fc1_b, o.fc2_w, o.fc2_b, m->s_x, Me, site_seed(seed, 102 + l * 8), site_seed(seed, 103 + l * 8), st));
    void* dh_mid = m->s_b;
    TS_TRY(layernorm_bwd(ctx, dt, m->s_x, b.h_mid, m->P + o.ln2_g, b.m2, b.r2, dh, dh_mid, m->G + o.ln2_g, m->G + o.ln2_b, (int)Me, d, st));
    TS_TRY(dense_bwd(m, b.ctx, d, dh_mid, d, o.sa.o_w, d, o.sa.o_b, m->s_ctx, nullptr, Me, st));
    const char* qkv = (const char*)b.qkv;
    char* dqkv = (char*)m->s_qkv;
    TS_TRY(attn_backward(m, qkv, 3 * d, qkv + m->E(d), qkv + m->E(2 * d), 3 * d, b.P, m->s_ctx, d, dqkv, 3 * d, dqkv + m->E(d), dqkv + m->E(2 * d), 3 * d,
                         {B, nh, hd, T, T, Tp}, scale, m->drop(c.attention_dropout), site_seed(seed, 100 + l * 8), st, b.ctx, 0));
    TS_TRY(dense_bwd(m, b.x1, d, dqkv, 3 * d, o.sa.qkv_w, 3 * d, o.sa.qkv_b, m->s_x, nullptr, Me, st));
    TS_TRY(layernorm_bwd(ctx, dt, m->s_x, b.h_in, m->P + o.ln1_g, b.m1, b.r1, dh_mid, m->s_a, m->G + o.ln1_g, m->G + o.ln1_b, (int)Me, d, st));
    return 0;
  }
  // ---- conv stem --------------------------------------------------------------------------------------------------
  void* dh0 = m->s_a;
  if (m->drop(c.dropout) > 0) TS_TRY(dropout_apply(ctx, dt, dh0, dh0, Me * d, c.dropout, site_seed(seed, 1), st));
  void* du2 = m->s_du;
  TS_TRY(gelu_bwd_rows(ctx, dt, dh0, T, nullptr, m->u2, du2, m->Rq2, B, T, d, st));
  TS_TRY(colsum_acc(ctx, dt, du2, d, (int)(B * m->Rq2), d, m->G + m->conv2_b, st));
  TS_TRY(GemmB(dt, TS_F32).A(m->a1, 1, 2 * d).B(du2, 1, d).C(m->G + m->conv2_w, d).mnk(3 * d, d, (int)(B * m->Rq2)).acc().run(ctx, st));
  TS_TRY(GemmB(dt, dt).A(du2, 0, d).B(m->W(m->conv2_w), 0, d).C(m->s_dcol, 3 * d).mnk((int)(B * m->Rq2), 3 * d, d).run(ctx, st));
  Col2imSrc col;
  col.dcol = m->s_dcol; col.rows_per_batch = m->Rq2; col.t_next = T; col.k = 3; col.s = 2; col.left = 0;
  void* du1 = m->s_du;  // du2 is dead once dcol exists
  TS_TRY(gelu_bwd_rows(ctx, dt, nullptr, 0, &col, m->u1, du1, m->Rq1, B, m->Tm, d, st));
  TS_TRY(colsum_acc(ctx, dt, du1, d, (int)(B * m->Rq1), d, m->G + m->conv1_b, st));
  TS_TRY(GemmB(dt, TS_F32).A(m->xT, 1, c.n_mels).B(du1, 1, d).C(m->G + m->conv1_w, d).mnk(3 * c.n_mels, d, (int)(B * m->Rq1)).acc().run(ctx, st));
  return 0;
}

}  // namespace ts

using namespace ts;

extern "C" {

int ts_whisper_create(ts_ctx* ctx_, const ts_whisper_config* cfg, int precision, ts_whisper** out) {
  Ctx* ctx = reinterpret_cast<Ctx*>(ctx_);
  if (!ctx || !cfg || !out) return TS_EINVAL;
  TS_REQUIRE(ctx, precision == TS_F32 || precision == TS_BF16, TS_EDTYPE, "whisper: precision must be TS_F32 or TS_BF16");
  TS_REQUIRE(ctx, cfg->vocab > 0 && cfg->start_token >= 0 && cfg->start_token < cfg->vocab, TS_EINVAL,
             "whisper: decoder start token %d outside the vocabulary of %d entries", cfg->start_token, cfg->vocab);
  Whisper* m = new Whisper();
  m->ctx = ctx; m->cfg = *cfg; m->prec = precision; m->esz = precision == TS_BF16 ? 2 : 4;
  m->fused_attn = precision == TS_BF16 && cfg->heads > 0 && cfg->d_model / cfg->heads == 64 && !getenv("TETHYS_UNFUSED_ATTENTION");
  build_params(m);
  // PositionalEncoding tables (W:55-64): float64 math, interleaved sin/cos, cast to fp32
  const int d = cfg->d_model;
  auto make_pe = [&](int len, float** dev) -> int {
    std::vector<float> h((size_t)len * d);
    for (int pos = 0; pos < len; ++pos)
      for (int i = 0; i < d; i += 2) {
        const double div = exp((double)i * -(log(10000.0) / (double)d));
        h[(size_t)pos * d + i] = (float)sin((double)pos * div);
        if (i + 1 < d) h[(size_t)pos * d + i + 1] = (float)cos((double)pos * div);
      }
    TS_CUDA_OK(ctx, cudaMalloc(dev, h.size() * sizeof(float)));
    TS_CUDA_OK(ctx, cudaMemcpy(*dev, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
  };
  if (make_pe(cfg->n_ctx, &m->pe_enc) || make_pe(cfg->max_target, &m->pe_dec)) { delete m; return TS_ECUDA; }
  if (!(getenv("TETHYS_NO_SIDE_WGRAD") && atoi(getenv("TETHYS_NO_SIDE_WGRAD")) != 0)) {
    if (cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      m->side = nullptr;   // no side stream: everything stays on the caller's stream
    }
    // second branch (cross K/V projections and their backward): measured neutral — 5.16 vs 5.11 ms (default preset), 5.42 vs 5.44 ms
    // (base): the full-size GEMMs hold every SM, so the decoder's small kernels queue behind them instead of running beside them.
    // Opt-in (TETHYS_SIDE_KV=1).
    if (getenv("TETHYS_SIDE_KV") && atoi(getenv("TETHYS_SIDE_KV")) != 0) {
      if (cudaStreamCreateWithFlags(&m->side2, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&m->ev_fork2, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&m->ev_join2, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        m->side2 = nullptr;
      }
    }
  }
  *out = reinterpret_cast<ts_whisper*>(m);
  return 0;
}
void ts_whisper_destroy(ts_whisper* h) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  if (!m) return;
  cudaFree(m->pe_enc); cudaFree(m->pe_dec);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  if (m->side) cudaStreamDestroy(m->side);
  if (m->ev_fork2) cudaEventDestroy(m->ev_fork2);
  if (m->ev_join2) cudaEventDestroy(m->ev_join2);
  if (m->side2) cudaStreamDestroy(m->side2);
  delete m;
}
int64_t ts_whisper_arena_elems(ts_whisper* h) { return reinterpret_cast<Whisper*>(h)->pt.n; }
int ts_whisper_num_params(ts_whisper* h) { return (int)reinterpret_cast<Whisper*>(h)->pt.defs.size(); }
int ts_whisper_param_info(ts_whisper* h, int i, char* name, int cap, int64_t* offset, int32_t* ndim, int64_t* shape4, int64_t* ld) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  if (i < 0 || i >= (int)m->pt.defs.size()) return TS_EINVAL;
  const ParamDef& d = m->pt.defs[i];
  if (name && cap > 0) { strncpy(name, d.name.c_str(), cap - 1); name[cap - 1] = 0; }
  if (offset) *offset = d.offset;
  if (ndim) *ndim = d.ndim;
  if (shape4) for (int j = 0; j < 4; ++j) shape4[j] = d.shape[j];
  if (ld) *ld = d.ld;
  return 0;
}
int ts_whisper_num_stages(ts_whisper* h) { return (int)reinterpret_cast<Whisper*>(h)->stage_end.size(); }
int64_t ts_whisper_stage_end(ts_whisper* h, int stage) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  if (stage < 0 || stage >= (int)m->stage_end.size()) return -1;
  return m->stage_end[stage];
}
int64_t ts_whisper_workspace_bytes(ts_whisper* h, int B, int Tm, int S) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  Whisper tmp = *m;
  Bump bp;
  if (plan(&tmp, B, Tm, S, bp)) return -1;
  return (int64_t)bp.off + 4096;
}
int ts_whisper_bind(ts_whisper* h, float* params, float* grads, void* params_lp, void* ws, int64_t ws_bytes) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  TS_REQUIRE(m->ctx, params && grads && ws, TS_EINVAL, "whisper_bind: null arena");
  TS_REQUIRE(m->ctx, m->prec != TS_BF16 || params_lp, TS_EINVAL, "whisper_bind: bf16 mode needs the bf16 parameter arena");
  m->P = params; m->G = grads; m->P16 = params_lp; m->ws = (char*)ws; m->ws_bytes = ws_bytes;
  m->planned = false;
  return 0;
}
int ts_whisper_sync_compute_weights(ts_whisper* h, void* stream) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  if (m->prec == TS_BF16) TS_TRY(cast_f32_to_bf16(m->ctx, m->P, m->P16, m->pt.n, (cudaStream_t)stream));
  return 0;
}
int ts_whisper_forward(ts_whisper* h, const float* feats, int B, int Tm, const int32_t* labels, int S, uint64_t seed, int training,
                       int compute_loss, void* stream) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  Ctx* ctx = m->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  TS_REQUIRE(ctx, m->P && m->ws, TS_EINVAL, "whisper_forward: call ts_whisper_bind first");
  TS_REQUIRE(ctx, B > 0 && feats && labels, TS_EINVAL, "whisper_forward: bad arguments");
  if (!m->planned || m->B != B || m->Tm != Tm || m->S != S) {
    Bump bp;
    bp.base = m->ws;
    TS_TRY(plan(m, B, Tm, S, bp));
    TS_REQUIRE(ctx, (long long)bp.off <= m->ws_bytes, TS_EINVAL, "whisper_forward: workspace too small (%lld < %lld bytes)",
               (long long)m->ws_bytes, (long long)bp.off);
    m->planned = true;
  }
  m->seed = seed; m->training = training; m->fwd_done = false; m->gen_ready = false;
  TS_TRY(whisper_forward(m, feats, labels, st));
  if (compute_loss) TS_TRY(whisper_loss(m, st));
  return 0;
}
// ---- greedy decoding: WhisperForConditionalGeneration.generate (W:636-709) -------------------------------------------
static int whisper_ensure_plan(Whisper* m, int B, int Tm, int S, const char* who) {
  Ctx* ctx = m->ctx;
  TS_REQUIRE(ctx, m->P && m->ws, TS_EINVAL, "%s: call ts_whisper_bind first", who);
  if (!m->planned || m->B != B || m->Tm != Tm || m->S != S) {
    Bump bp;
    bp.base = m->ws;
    TS_TRY(plan(m, B, Tm, S, bp));
    TS_REQUIRE(ctx, (long long)bp.off <= m->ws_bytes, TS_EINVAL, "%s: workspace too small (%lld < %lld bytes)", who,
               (long long)m->ws_bytes, (long long)bp.off);
    m->planned = true;
  }
  return 0;
}
int ts_whisper_encode(ts_whisper* h, const float* feats, int B, int Tm, int max_len, void* stream) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  Ctx* ctx = m->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  TS_REQUIRE(ctx, B > 0 && feats && max_len >= 1, TS_EINVAL, "whisper_encode: bad arguments");
  TS_TRY(whisper_ensure_plan(m, B, Tm, max_len < 2 ? 2 : max_len, "whisper_encode"));
  m->seed = 0; m->training = 0; m->fwd_done = false; m->gen_ready = false;
  TS_TRY(whisper_encoder_forward(m, feats, st));
  // the cross-attention keys / values depend on the encoder output only: project them once for all decode steps
  const ts_whisper_config& c = m->cfg;
  const int d = c.d_model;
  const long long Me = (long long)B * m->T;
  for (int l = 0; l < c.dec_layers; ++l)
    TS_TRY(GemmB(m->prec, m->prec).A(m->enc_out, 0, d).B(m->W(m->DL[l].ca.kv_w), 1, 2 * d).C(m->DB[l].kv, 2 * d)
               .bias(m->P + m->DL[l].ca.kv_b).mnk((int)Me, 2 * d, d).run(ctx, st));
  m->gen_ready = true;
  return 0;
}
int ts_whisper_decode_step(ts_whisper* h, int32_t* tokens, int64_t ld_tok, int len, void* stream) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  Ctx* ctx = m->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  TS_REQUIRE(ctx, m->planned && m->gen_ready, TS_EINVAL, "whisper_decode_step: call ts_whisper_encode first");
  TS_REQUIRE(ctx, tokens && len >= 1 && len <= m->S && ld_tok >= len, TS_EINVAL,
             "whisper_decode_step: len %d outside [1, %d] (the max_len given to ts_whisper_encode) or token stride too small", len, m->S);
  TS_TRY(whisper_decoder_forward(m, tokens, ld_tok, len, true, true, st));
  if (m->prec == TS_F32) ts::launch_k(argmax_rows_kernel<float>, m->B, 256, 0, st, (const float*)m->logits, m->Vp, m->cfg.vocab, tokens + (len - 1), ld_tok);
  else ts::launch_k(argmax_rows_kernel<bf16>, m->B, 256, 0, st, (const bf16*)m->logits, m->Vp, m->cfg.vocab, tokens + (len - 1), ld_tok);
  TS_LAUNCH_OK(ctx);
  return 0;
}
int ts_whisper_backward(ts_whisper* h, int stage_from, int stage_to, void* stream) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  cudaStream_t st = (cudaStream_t)stream;
  TS_REQUIRE(m->ctx, m->fwd_done, TS_EINVAL, "whisper_backward: run forward with compute_loss=1 first");
  const int ns = (int)m->stage_end.size();
  if (stage_from <= 0) TS_TRY(fill_zero(m->ctx, m->G, 4ll * m->pt.n, st));
  // stage map: 0 lm_head | 1..Ld decoder layers | Ld+1 embedding + encoder final LN | Ld+2..Ld+1+Le encoder layers | last conv stem
  const int total = m->cfg.dec_layers + m->cfg.enc_layers + 3;
  for (int s = std::max(0, stage_from); s <= std::min(total - 1, stage_to); ++s) TS_TRY(whisper_backward_stage(m, s, st));
  (void)ns;
  return 0;
}
int ts_whisper_step(ts_whisper* h, const float* feats, int B, int Tm, const int32_t* labels, int S, const ts_step_args* a, void* stream) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  if (!m) return TS_EINVAL;
  Ctx* ctx = m->ctx;
  TS_REQUIRE(ctx, a, TS_EINVAL, "whisper_step: args are required");
  TS_TRY(ts_whisper_forward(h, feats, B, Tm, labels, S, a->seed, a->dropout ? 1 : 0, 1, stream));
  TS_TRY(ts_whisper_backward(h, 0, 1 << 20, stream));
  // W:829-836: raw SUM of the replicas' gradients (no 1/N), Adam without clipping unless the caller asks for it
  return step_reduce_update(ctx, m->P, m->G, m->P16, m->pt.n, m->scalars, /*loss_mean_over_replicas=*/false, a, (cudaStream_t)stream);
}
int ts_whisper_get_buffer(ts_whisper* h, const char* name, void** ptr, int32_t* dtype, int32_t* ndim, int64_t* shape4) {
  Whisper* m = reinterpret_cast<Whisper*>(h);
  if (!m->planned) return set_err(m->ctx, TS_EINVAL, "whisper_get_buffer: no plan yet");
  const std::string s(name);
  auto set = [&](void* p, int dt, int nd, long long a, long long b, long long d3, long long d4) {
    *ptr = p; *dtype = dt; *ndim = nd; shape4[0] = a; shape4[1] = b; shape4[2] = d3; shape4[3] = d4; return 0; };
  const long long B = m->B, d = m->cfg.d_model;
  if (s == "scalars") return set(m->scalars, TS_F32, 1, 4, 1, 1, 1);
  if (s == "encoder_last_hidden_state") return set(m->enc_out, m->prec, 3, B, m->T, d, 1);
  if (s == "last_hidden_state") return set(m->dec_out, m->prec, 3, B, m->S, d, 1);
  if (s == "next_token_logits") return set(m->logits, m->prec, 2, B, m->Vp, 1, 1);   // after ts_whisper_decode_step
  if (s == "logits") return set(m->logits, m->prec, 3, B, m->S, m->Vp, 1);   // padded last dim; valid columns [0, vocab)
  if (s == "decoder_self_attn_probs0") {
    if (m->fused_attn) return set_err(m->ctx, TS_EUNSUPPORTED, "probabilities are not materialised by the fused attention kernels; read decoder_self_attn_stats0");
    return set(m->DB[0].P, m->prec, 4, B, m->cfg.heads, m->S, m->Sp);
  }
  if (s == "decoder_self_attn_stats0") {  // [B, heads, S, 2] = (row max, log row-sum) of the fused kernel
    if (!m->fused_attn) return set_err(m->ctx, TS_EUNSUPPORTED, "row statistics exist only with the fused attention kernels");
    return set(m->DB[0].P, TS_F32, 4, B, m->cfg.heads, m->S, 2);
  }
  return set_err(m->ctx, TS_EINVAL, "whisper_get_buffer: unknown buffer '%s'", name);
}

}  // extern "C"
