"""CUDA-graph replay of the whole train step (runtime.GraphedTrainStep): same parameters as the eager step after the same
number of steps, fresh dropout masks on every replay (device-resident salt), Adam's step count advanced on the device."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed=0, precision="bf16"):
    from tethys_speech_b200 import wav2vec2 as W
    from tethys_speech_b200.runtime import Adam

    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("tiny"), precision=precision, device=0, seed=seed)
    opt = Adam(learning_rate=1e-4, epsilon=1e-8, clipnorm=1.0)
    g = torch.Generator().manual_seed(3)
    wave = torch.randn(2, 6400, generator=g).cuda()
    T = model.num_frames(6400)
    neg = model._sample_negative_indices(T, 2)[:, 0, :].contiguous()
    return W, model, opt, wave, neg


def test_graphed_step_matches_eager_without_dropout():
    from tethys_speech_b200.runtime import GraphedTrainStep

    W, m1, o1, wave, neg = _setup()
    _, m2, o2, _, _ = _setup()
    fn1 = lambda batch, aux: W.train_step(m1, batch, o1, neg_indices=aux["neg"], dropout=False)
    graphed = GraphedTrainStep(fn1, m1, o1, (wave, None), {"neg": neg}, warmup=2)
    for _ in range(3):
        loss_g = graphed((wave, None), {"neg": neg})
    for _ in range(3):                                   # the 2 warm-up passes leave no trace: 3 replays == 3 eager steps
        loss_e = W.train_step(m2, (wave, None), o2, neg_indices=neg, dropout=False)
    torch.cuda.synchronize()
    assert o1.iterations == o2.iterations == 3
    # bf16 training is sensitive to the summation order of the split-K atomics (run-to-run differences of ~1e-3 in the loss
    # after a few steps, eager or graphed alike): compare loosely, the step COUNT (bias correction) is what must be exact
    assert abs(float(loss_g) - float(loss_e)) < 2e-2 * abs(float(loss_e))
    p1, p2 = m1._prog.params, m2._prog.params
    assert float((p1 - p2).norm() / p2.norm()) < 2e-3


def test_graphed_step_draws_new_dropout_masks_each_replay():
    from tethys_speech_b200.runtime import GraphedTrainStep

    W, m, o, wave, neg = _setup(1)
    o.learning_rate = 0.0                                  # freeze the weights: only the dropout mask can change the loss
    graphed = GraphedTrainStep(lambda batch, aux: W.train_step(m, batch, o, neg_indices=aux["neg"]), m, o, (wave, None),
                               {"neg": neg}, warmup=1)
    losses = [float(graphed((wave, None), {"neg": neg})) for _ in range(4)]
    assert len(set(losses)) == 4, losses
    m._prog.ctx.watchdog()


def test_graphed_step_fp32_step_count_is_exact():
    """fp32 mode has no split-K atomics: after the same number of steps the graphed and the eager parameters agree to
    rounding, which pins the device-side Adam step count (a bias-correction step off by one changes the update by >10 %)."""
    from tethys_speech_b200.runtime import GraphedTrainStep

    W, m1, o1, wave, neg = _setup(2, "fp32")
    _, m2, o2, _, _ = _setup(2, "fp32")
    p0 = m2._prog.params.clone()
    graphed = GraphedTrainStep(lambda batch, aux: W.train_step(m1, batch, o1, neg_indices=aux["neg"], dropout=False), m1, o1,
                               (wave, None), {"neg": neg}, warmup=1)
    assert float((m1._prog.params - p0).abs().max()) == 0.0      # building the graph (1 warm-up pass) did not train
    for _ in range(3):
        graphed((wave, None), {"neg": neg})
    for _ in range(3):
        W.train_step(m2, (wave, None), o2, neg_indices=neg, dropout=False)
    torch.cuda.synchronize()
    upd = (m2._prog.params - p0).norm()
    assert float((m1._prog.params - m2._prog.params).norm() / upd) < 1e-2
