"""TEST INFRASTRUCTURE ONLY — CPU restatement of the TensorFlow/Keras 2.10 op semantics that
tethys-speech's train step relies on.  Nothing under oracle/ may be imported
by the product path (tethys_speech_b200/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs use it, and only as the checker / reported baseline.

PINNING STATUS: the reference ships no tests, golden vectors or seeds and TensorFlow
(Dockerfile:1 pins nvcr.io/nvidia/tensorflow:22.12-tf2-py3 = TF 2.10 / Keras 2.10) is not installable here.
What pins this oracle instead:
  (a) STRUCTURE — pinned against the reference's own code: the UNMODIFIED /root/reference/speech_jobs scripts are
      imported on a tf-on-torch stand-in (oracle/tf_shim.py) and their model classes, gradients and step functions
      (distributed_train_step of V and W at 1 and 2 replicas, train_step of VS and WS) agree with this oracle to 1e-10
      in float64 (tests/test_reference_pinning.py; golden vectors of those runs: tests/golden/ref_*.npz);
  (b) OP SEMANTICS — still restated (here AND, independently, in tf_shim.py) from the published TF behaviour, and
      cross-checked by a second numpy-fp64 restatement (oracle/np_forward.py), hand-computed micro-cases
      (tests/test_oracle_tf_semantics.py) and finite differences. No real TensorFlow run backs them: in that sense
      parity with TF's kernels stays unpinned.

Every function cites the reference call site it follows (W = speech_jobs/whisper_dist.py,
V = speech_jobs/wav2vec2_dist.py).
"""
import contextlib
import math

import torch
import torch.nn.functional as F

# ---- bf16-storage emulation (tests only: the per-tensor error budget of the bf16 CUDA path) -------------------
# With BF16_STORAGE on, every op below keeps its arithmetic in the tensor's dtype (fp64 in the tests) but rounds what a
# mixed-precision run would STORE in bfloat16: the kernel/weight operand of every Dense/Conv (the bf16 compute copy of the
# fp32 master weights), every activation an op returns, and — through the backward of q() — every activation gradient.
# Accumulation, statistics, losses and weight gradients stay unrounded (fp32 on the GPU). It does not model the CUDA
# kernels (different rounding points, fused epilogues); it measures how far bf16 storage ALONE moves each tensor away
# from the fp64 result, which is what a tolerance for the bf16 path has to allow for.
BF16_STORAGE = False


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def q(x):
    """Identity unless bf16-storage emulation is on: then round-to-nearest-even to bfloat16 (value and gradient)."""
    return _RoundBF16.apply(x) if BF16_STORAGE else x


def qw(w):
    """bf16 compute copy of a master weight (value rounded; the weight gradient itself stays unrounded)."""
    return w + (w.to(torch.bfloat16).to(w.dtype) - w).detach() if BF16_STORAGE else w


@contextlib.contextmanager
def bf16_storage():
    global BF16_STORAGE
    old, BF16_STORAGE = BF16_STORAGE, True
    try:
        yield
    finally:
        BF16_STORAGE = old


def same_pad(t_in: int, k: int, s: int):
    """TF 'SAME' padding (A-1): T_out = ceil(T/s); pad_total = max((T_out-1)*s + k - T, 0);
    pad_left = pad_total // 2, the extra element goes to the RIGHT (unlike torch's symmetric pad)."""
    t_out = -(-t_in // s)
    pad_total = max((t_out - 1) * s + k - t_in, 0)
    left = pad_total // 2
    return t_out, left, pad_total - left


def conv1d_same(x, kernel, stride=1, groups=1, bias=None):
    """tf.keras.layers.Conv1D(padding='same') — W:311-312, V:240-247, V:257-264, V:271-277.
    x [B,T,Cin] channels-last; kernel [k, Cin/groups, Cout] (Keras layout); cross-correlation."""
    k = kernel.shape[0]
    _, left, right = same_pad(x.shape[1], k, stride)
    xt = F.pad(x.transpose(1, 2), (left, right))          # [B,Cin,T+pad]
    w = qw(kernel).permute(2, 1, 0)                       # [Cout, Cin/groups, k]
    y = F.conv1d(xt, w, bias=bias, stride=stride, groups=groups)
    return q(y.transpose(1, 2))


def dense(x, kernel, bias=None):
    """tf.keras.layers.Dense (A-3): y = x @ W[in,out] + b."""
    y = x @ qw(kernel)
    return q(y if bias is None else y + bias)


def layer_norm(x, gamma, beta, eps=1e-5):
    """tf.keras.layers.LayerNormalization(epsilon=1e-5) (A-4): last axis, biased variance."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return q((x - mu) * torch.rsqrt(var + eps) * gamma + beta)


def gelu(x):
    """Exact-erf GELU (A-5) — V:132-136 and tf.keras.activations.gelu default (W:195, W:333)."""
    return q(0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0))))


def group_norm(x, gamma, beta, groups, eps=1e-5):
    """GroupNormalization.call — V:167-196.  x [B,T,C]; reshape [B,T,G,C/G], transpose to
    [B,T,C/G,G], tf.nn.moments over axes (1,2) (= time x channels-in-group, biased variance, A-15),
    normalise, then per-channel gamma/beta."""
    b, t, c = x.shape
    xg = x.reshape(b, t, groups, c // groups)
    mu = xg.mean(dim=(1, 3), keepdim=True)
    var = ((xg - mu) ** 2).mean(dim=(1, 3), keepdim=True)
    xn = (xg - mu) / torch.sqrt(var + eps)
    return q(gamma * xn.reshape(b, t, c) + beta)


def softmax_xent_sparse(logits, labels):
    """tf.nn.sparse_softmax_cross_entropy_with_logits / SparseCategoricalCrossentropy(from_logits,
    reduction=NONE) (A-7): per-element loss."""
    lse = torch.logsumexp(logits, dim=-1)
    picked = torch.gather(logits, -1, labels.long().unsqueeze(-1)).squeeze(-1)
    return lse - picked


def sinusoid_pe(max_len, d_model, dtype=torch.float32):
    """PositionalEncoding.__init__ — W:55-64: float64 numpy table, even=sin, odd=cos, cast to fp32."""
    import numpy as np

    pe = np.zeros((max_len, d_model))
    position = np.arange(0, max_len)[:, np.newaxis]
    div_term = np.exp(np.arange(0, d_model, 2) * -(np.log(10000.0) / d_model))
    pe[:, 0::2] = np.sin(position * div_term)
    pe[:, 1::2] = np.cos(position * div_term)
    return torch.from_numpy(pe.astype(np.float32)).to(dtype)


# ---- optimiser / clipping (A-11, A-12) ---------------------------------------------------------
def clip_by_global_norm(grads, clip_norm=1.0):
    """tf.clip_by_global_norm — V:1243: n = sqrt(sum ||g||^2); g * clip/max(n, clip)."""
    n = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    scale = clip_norm / torch.clamp(n, min=clip_norm)
    return [g * scale for g in grads], n


def clip_by_norm_each(grads, clip_norm=1.0):
    """Keras optimizer clipnorm=1.0 (V:1274): per variable g * clip/max(||g||, clip)."""
    out = []
    for g in grads:
        n = torch.sqrt((g.double() ** 2).sum()).to(g.dtype)
        out.append(g * (clip_norm / torch.clamp(n, min=clip_norm)))
    return out


def keras_adam_step(params, grads, ms, vs, t, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """tf.keras.optimizers.Adam in Keras 2.10 (legacy OptimizerV2, A-12), in place:
         m += (g-m)(1-b1); v += (g^2-v)(1-b2); p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)
    eps sits OUTSIDE the bias correction (differs from torch.optim.Adam). t starts at 1."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    with torch.no_grad():
        for p, g, m, v in zip(params, grads, ms, vs):
            m.add_((g - m) * (1.0 - beta1))
            v.add_((g * g - v) * (1.0 - beta2))
            p.sub_(lr_t * m / (torch.sqrt(v) + eps))


def glorot_uniform(gen, shape, fan_in, fan_out, dtype=torch.float32):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit).to(dtype)
