"""Checkpoint save / restore (reference save sites V:1286-1288, V:1341, V:1362, W:956, W:1025).
CPU: the file container and the arena <-> variable mapping (incl. column slices of fused q/k/v blocks), error behaviour.
GPU: a restored run continues bit-identically to the uninterrupted one (parameters, Adam moments, step count, dropout seed)."""
import json
import os

import numpy as np
import pytest
import torch

from tethys_speech_b200 import checkpoint as CK


def _fake_info():
    # one dense vector, one dense matrix, three column slices of a fused [4, 3*5 (+1 pad)] block, one 3-d kernel
    return {
        "ln.gamma": (0, (7,), 7),
        "dense.kernel": (64, (3, 4), 12),
        "attn.q.kernel": (128, (4, 5), 16),
        "attn.k.kernel": (133, (4, 5), 16),
        "attn.v.kernel": (138, (4, 5), 16),
        "conv.kernel": (256, (2, 3, 4), 24),
    }


def test_gather_scatter_roundtrip_with_fused_blocks():
    info = _fake_info()
    rng = np.random.default_rng(0)
    arena = rng.standard_normal(320).astype(np.float32)
    vars_ = CK.gather_variables(arena, info)
    assert vars_["attn.k.kernel"].shape == (4, 5)
    # column slice: row r of k starts at 133 + 16 r
    for r in range(4):
        np.testing.assert_array_equal(vars_["attn.k.kernel"][r], arena[133 + 16 * r:138 + 16 * r])
    np.testing.assert_array_equal(vars_["conv.kernel"].reshape(-1), arena[256:280])
    other = np.full(320, -7.0, np.float32)
    done = CK.scatter_variables(other, info, vars_)
    assert sorted(done) == sorted(info)
    again = CK.gather_variables(other, info)
    for k in info:
        np.testing.assert_array_equal(again[k], vars_[k])
    # bytes outside any variable (alignment gaps, the pad column of the fused block) are left alone
    assert other[7] == -7.0 and other[128 + 15] == -7.0 and other[300] == -7.0


def test_scatter_errors():
    info = _fake_info()
    arena = np.zeros(320, np.float32)
    vars_ = CK.gather_variables(arena, info)
    missing = {k: v for k, v in vars_.items() if k != "ln.gamma"}
    with pytest.raises(KeyError):
        CK.scatter_variables(arena, info, missing)
    assert "ln.gamma" not in CK.scatter_variables(arena, info, missing, strict=False)
    bad = dict(vars_)
    bad["dense.kernel"] = np.zeros((4, 3), np.float32)
    with pytest.raises(ValueError):
        CK.scatter_variables(arena, info, bad)
    extra = dict(vars_)
    extra["nope"] = np.zeros(1, np.float32)
    with pytest.raises(KeyError):
        CK.scatter_variables(arena, info, extra)


def test_file_container_roundtrip_alignment_and_truncation(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {"model/a": rng.standard_normal((3, 5)).astype(np.float32),
               "model/b": rng.standard_normal(1).astype(np.float32),
               "optimizer/m/a": np.zeros((0,), np.float32),
               "ids": np.arange(11, dtype=np.int64)}
    path = str(tmp_path / "x.tsckpt")
    CK.write_file(path, tensors, {"hello": 1})
    assert not os.path.exists(path + ".tmp")
    meta, back = CK.read_file(path)
    assert meta == {"hello": 1}
    assert list(back) == list(tensors)
    for k in tensors:
        assert back[k].dtype == tensors[k].dtype and back[k].shape == tensors[k].shape
        np.testing.assert_array_equal(back[k], tensors[k])
    # header is self-describing JSON and every array starts 64-byte aligned in the file
    raw = open(path, "rb").read()
    assert raw[:8] == CK.MAGIC
    hlen = int.from_bytes(raw[8:16], "little")
    hdr = json.loads(raw[16:16 + hlen])
    assert (16 + hlen) % 64 == 0 and all(d["offset"] % 64 == 0 for d in hdr["tensors"].values())
    _, only = CK.read_file(path, keys=lambda k: k.startswith("model/"))
    assert sorted(only) == ["model/a", "model/b"]
    with open(path, "wb") as f:
        f.write(raw[:-8])
    with pytest.raises(ValueError):
        CK.read_file(path)
    with open(path, "wb") as f:
        f.write(b"not a checkpoint at all")
    with pytest.raises(ValueError):
        CK.read_file(path)


def test_latest_checkpoint_and_tf_keys(tmp_path):
    assert CK.latest_checkpoint(str(tmp_path)) is None
    for n in (1, 2, 10):
        CK.write_file(str(tmp_path / f"model_step-{n}.tsckpt"), {}, {})
    assert CK.latest_checkpoint(str(tmp_path)).endswith("model_step-10.tsckpt")
    assert CK.tf_object_key("encoder.layers.0.attention.q_proj.kernel") == \
        "model/wav2vec2/encoder/layers/0/attention/q_proj/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    assert CK.tf_object_key("fe.conv2.gn.gamma") == \
        "model/wav2vec2/feature_extractor/conv_layers/2/layer_with_weights-1/gamma/.ATTRIBUTES/VARIABLE_VALUE"
    assert CK.tf_object_key("decoder.layers.1.encoder_attn.k_proj.bias", kind="whisper") == \
        "model/model/decoder/layers/1/encoder_attn/k_proj/bias/.ATTRIBUTES/VARIABLE_VALUE"
    assert CK.tf_slot_key("lm_head.kernel", "m", kind="whisper") == \
        "model/lm_head/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"
    # files without a save counter (plain save(path, ...)) are found too, ranked by modification time below numbered ones
    other = tmp_path / "plain"
    other.mkdir()
    CK.write_file(str(other / "model_epoch_1.tsckpt"), {}, {})
    os.utime(str(other / "model_epoch_1.tsckpt"), (1, 1))
    CK.write_file(str(other / "model_step_50.tsckpt"), {}, {})
    assert CK.latest_checkpoint(str(other)).endswith("model_step_50.tsckpt")
    CK.write_file(str(other / "model_step_10-1.tsckpt"), {}, {})
    assert CK.latest_checkpoint(str(other)).endswith("model_step_10-1.tsckpt")


@pytest.mark.gpu
def test_resume_from_a_directory_written_by_the_train_loop(tmp_path, monkeypatch):
    """train_wav2vec2 writes its epoch checkpoint through a Checkpoint object (TF-style `<name>-<n>` files); `--resume <dir>` must
    find it and continue at the saved step instead of restarting from scratch."""
    from tethys_speech_b200 import train as TR
    from tethys_speech_b200.runtime import Strategy

    monkeypatch.setattr(TR, "WORKSPACE", str(tmp_path))
    st = Strategy()
    TR.train_wav2vec2(st, model_size="tiny", batch_size=2, num_batches=3, audio_length=6400, precision="fp32", cuda_graph=False)
    d = tmp_path / "checkpoints"
    files = sorted(os.listdir(d))
    assert files == ["model_epoch_1-1.tsckpt"], files
    from tethys_speech_b200 import wav2vec2 as W
    from tethys_speech_b200.runtime import Adam

    model = W.create_full_model("pretraining", "tiny", precision="fp32", device=0)
    opt = Adam(learning_rate=3e-5, epsilon=1e-8, clipnorm=1.0)
    assert TR._maybe_resume(str(d), model, opt) == 3 and opt.iterations == 3


# ---------------------------------------------------------------------------------------------------------------------
def _w2v(seed, precision):
    from tethys_speech_b200 import wav2vec2 as W
    from tethys_speech_b200.runtime import Adam

    model = W.Wav2Vec2ForPreTraining(W.Wav2Vec2Config("tiny"), precision=precision, device=0, seed=seed)
    opt = Adam(learning_rate=1e-3, epsilon=1e-8, clipnorm=1.0)
    return W, model, opt


@pytest.mark.gpu
def test_restore_continues_the_run_w2v_fp32(tmp_path):
    """Steps 3-4 after a restore must reproduce steps 3-4 of the straight run with dropout on — which needs parameters, m, v,
    the iteration count and the dropout step seed all restored. fp32 mode differs between two executions only by the order
    of the atomic adds in the bias / norm gradient reductions, hence the 1e-6 bars (a wrong dropout mask or a bias-correction
    step off by one changes the loss in the second digit)."""
    g = torch.Generator().manual_seed(5)
    wave = torch.randn(2, 6400, generator=g).cuda()
    W, m1, o1 = _w2v(3, "fp32")
    T = m1.num_frames(6400)
    neg = m1._sample_negative_indices(T, 2)[:, 0, :].contiguous()
    for _ in range(2):
        W.train_step(m1, (wave, None), o1, neg_indices=neg)
    ck = CK.Checkpoint(model=m1, optimizer=o1)
    path = ck.save(str(tmp_path / "ckpt" / "model_step"))
    assert path.endswith("model_step-1.tsckpt") and CK.latest_checkpoint(str(tmp_path / "ckpt")) == path
    straight = [float(W.train_step(m1, (wave, None), o1, neg_indices=neg)) for _ in range(2)]

    _, m2, o2 = _w2v(99, "fp32")                      # different init: everything must come from the file
    meta = CK.Checkpoint(model=m2, optimizer=o2).restore(path)
    assert meta["optimizer"]["iterations"] == 2 and o2.iterations == 2 and meta["kind"] == "Wav2Vec2ForPreTraining"
    resumed = [float(W.train_step(m2, (wave, None), o2, neg_indices=neg)) for _ in range(2)]
    assert all(abs(a - b) <= 1e-6 * abs(b) for a, b in zip(resumed, straight)), (resumed, straight)
    s1, s2 = o1._bind(m1), o2._bind(m2)
    for a, b in ((m1._prog.params, m2._prog.params), (s1["m"], s2["m"]), (s1["v"], s2["v"])):
        assert float((a - b).norm() / b.norm()) < 1e-6
    # every variable the model lists is in the file under its Keras path, with the fused q/k/v blocks split out again
    _, tensors = CK.read_file(path)
    assert {k[6:] for k in tensors if k.startswith("model/")} == set(m1.variable_names)
    assert tensors["model/encoder.layers.0.attention.q_proj.kernel"].shape == (m1.config.hidden_size, m1.config.hidden_size)
    # the same tensors under the names the reference's tf.train.Checkpoint(model=, optimizer=) gives them
    named = CK.tf_named_tensors(path)
    assert len(named) == 3 * len(m1.variable_names) + 1 and int(named["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"]) == 2
    k = "model/wav2vec2/encoder/layers/0/attention/q_proj/kernel"
    assert named[k + "/.ATTRIBUTES/VARIABLE_VALUE"].shape == (m1.config.hidden_size, m1.config.hidden_size)
    assert named[k + "/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE"].shape == (m1.config.hidden_size, m1.config.hidden_size)
    m1._prog.ctx.watchdog()


@pytest.mark.gpu
def test_save_weights_load_weights_whisper_bf16(tmp_path):
    """model.save_weights (W:1025) / load_weights: variables only; the bf16 compute copy is refreshed after a load (same loss)."""
    from tethys_speech_b200 import whisper as WH

    cfg = WH.WhisperConfig()
    cfg.d_model, cfg.encoder_layers, cfg.decoder_layers, cfg.d_ff = 128, 1, 1, 256
    cfg.encoder_attention_heads = cfg.decoder_attention_heads = 2
    cfg.vocab_size, cfg.decoder_start_token_id = 512, 3
    m1 = WH.WhisperForConditionalGeneration(cfg, precision="bf16", device=0, seed=1)
    m2 = WH.WhisperForConditionalGeneration(cfg, precision="bf16", device=0, seed=2)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, cfg.n_mels, 200, generator=g).cuda()
    lab = torch.randint(3, 100, (2, 12), generator=g).to(torch.int32).cuda()
    l1 = float(m1(x, labels=lab, training=True, dropout=False)["loss"])
    l2 = float(m2(x, labels=lab, training=True, dropout=False)["loss"])
    assert abs(l1 - l2) > 1e-4 * l1
    path = str(tmp_path / "w.tsckpt")
    m1.save_weights(path)
    meta = m2.load_weights(path)
    assert "optimizer" not in meta
    assert abs(float(m2(x, labels=lab, training=True, dropout=False)["loss"]) - l1) <= 1e-6 * l1   # loss sum = atomic adds
    w1, w2 = m1.get_weights(), m2.get_weights()
    assert all(torch.equal(w1[k], w2[k]) for k in w1)
    from tethys_speech_b200.runtime import Adam
    with pytest.raises(KeyError):
        CK.restore(path, m2, Adam())                  # asks for optimizer state the file does not hold


def test_tf_object_keys_follow_the_reference_object_graph():
    """checkpoint.tf_object_key against the reference's OWN model objects: the unmodified constructors of wav2vec2_single.py /
    whisper_dist.py are run on the TF shim, their attribute graph is walked breadth first with TensorFlow 2.10's trackable
    naming rules (oracle/tf_object_graph.py) and every variable's first path must be the key the product maps that variable to.
    (Needs /root/reference; skipped where it is absent.)"""
    import torch

    from oracle import ref_runner as R
    from oracle import tf_object_graph as G

    if not R.available():
        pytest.skip("/root/reference is not present on this machine")
    ref = R.load("wav2vec2_single")
    wave = torch.randn(2, 3200, dtype=torch.float64)
    labels = torch.tensor([3, 7], dtype=torch.int32)
    for model_type, head in (("pretraining", "pretraining"), ("asr", "ctc"), ("classification", "classification")):
        model = R.build_w2v(ref, "tiny", wave, model_type=model_type, labels=labels)
        vm = R.w2v_variable_map(model, head=head)
        keys = G.variable_keys({"model": model})
        assert set(keys) == {id(v) for v in vm.values()}, model_type
        for name, var in vm.items():
            assert CK.tf_object_key(name, kind="wav2vec2") == keys[id(var)], (model_type, name)
    wref = R.load("whisper_dist")

    def small(c):
        c.d_model, c.d_ff, c.encoder_layers, c.decoder_layers = 64, 128, 2, 2
        c.encoder_attention_heads = c.decoder_attention_heads = 2
        c.vocab_size, c.n_mels, c.n_ctx, c.decoder_start_token_id = 203, 16, 64, 200

    wm = R.build_whisper(wref, small, torch.randn(1, 16, 100, dtype=torch.float64), torch.randint(0, 100, (1, 8), dtype=torch.int32))
    vm = R.whisper_variable_map(wm)
    keys = G.variable_keys({"model": wm})
    assert set(keys) == {id(v) for v in vm.values()}
    for name, var in vm.items():
        assert CK.tf_object_key(name, kind="whisper") == keys[id(var)], name


def test_export_tf_names_and_tf_named_tensors(tmp_path):
    """export_tf_names (what save() records as meta["tf_keys"]) and the TF-keyed view of a file, on a stub model (no GPU)."""
    import numpy as np

    class WhisperForConditionalGeneration:                       # the class name selects the rule table
        variable_names = ["encoder.conv1.kernel", "decoder.embed_tokens.embeddings", "lm_head.kernel"]

    class Wav2Vec2ForCTC:
        variable_names = ["fe.conv0.kernel", "fe.conv0.gn.beta", "fe.pos_conv.bias", "fe.layer_norm.gamma", "quantizer.codevectors",
                          "project_hid.dense.kernel", "lm_head.bias"]

    t = CK.export_tf_names(WhisperForConditionalGeneration(), with_optimizer=True)
    assert t["encoder.conv1.kernel"] == "model/model/encoder/conv1/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    assert t["lm_head.kernel"] == "model/lm_head/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    assert t["optimizer/v/decoder.embed_tokens.embeddings"] == \
        "model/model/decoder/embed_tokens/embeddings/.OPTIMIZER_SLOT/optimizer/v/.ATTRIBUTES/VARIABLE_VALUE"
    assert t["optimizer/iterations"] == "optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"
    w = CK.export_tf_names(Wav2Vec2ForCTC())
    assert w["fe.conv0.gn.beta"] == "model/wav2vec2/feature_extractor/conv_layers/0/layer_with_weights-1/beta/.ATTRIBUTES/VARIABLE_VALUE"
    assert w["fe.pos_conv.bias"] == "model/wav2vec2/feature_extractor/pos_conv_embed/bias/.ATTRIBUTES/VARIABLE_VALUE"
    assert w["fe.layer_norm.gamma"] == "model/wav2vec2/feature_extractor/layer_norm/gamma/.ATTRIBUTES/VARIABLE_VALUE"
    assert w["project_hid.dense.kernel"] == "model/wav2vec2/project_hid/dense/kernel/.ATTRIBUTES/VARIABLE_VALUE"
    assert w["lm_head.bias"] == "model/lm_head/bias/.ATTRIBUTES/VARIABLE_VALUE"
    assert len(set(w.values())) == len(w)
    # a file whose meta carries tf_keys reads back keyed by TF names
    path = str(tmp_path / "m-1.tsckpt")
    tensors = {"model/lm_head.bias": np.arange(4, dtype=np.float32), "optimizer/m/lm_head.bias": np.ones(4, np.float32)}
    meta = {"optimizer": {"iterations": 7}, "tf_keys": {"model/lm_head.bias": w["lm_head.bias"],
                                                        "optimizer/m/lm_head.bias": CK.tf_slot_key("lm_head.bias", "m")}}
    CK.write_file(path, tensors, meta)
    named = CK.tf_named_tensors(path)
    assert np.array_equal(named["model/lm_head/bias/.ATTRIBUTES/VARIABLE_VALUE"], np.arange(4, dtype=np.float32))
    assert int(named["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"]) == 7
    assert "model/lm_head/bias/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE" in named
    CK.write_file(path, tensors, {})
    with pytest.raises(KeyError):
        CK.tf_named_tensors(path)


def test_save_restore_host_logic_on_a_stub_program(tmp_path, monkeypatch):
    """save() / restore() end to end on CPU arenas behind a stub of the native program (no kernels involved in a checkpoint):
    fused-block column slices, Adam slots, iteration count, dropout step seed, device step state, the recorded TF keys."""
    import ctypes as C

    import torch

    from tethys_speech_b200 import runtime as RT

    calls = {}

    class Lib:
        def ts_step_state_get(self, h, salt, step):
            salt._obj.value, step._obj.value = 11, 2
            return 0

        def ts_step_state_set(self, h, salt, step, stream):
            calls["state_set"] = (int(salt), int(step))
            return 0

    class Ctx:
        h, lib = None, Lib()

        def check(self, rc):
            assert rc == 0

    class Prog:
        def __init__(self, fill):
            # a dense 3-D conv kernel, a [4, 4] column slice of a fused q/k/v block (row stride 12), a bias
            self.info = {"fe.conv0.kernel": (0, (10, 1, 4), 40), "encoder.layers.0.attention.q_proj.kernel": (64, (4, 4), 12),
                         "encoder.layers.0.attention.k_proj.kernel": (68, (4, 4), 12), "lm_head.bias": (128, (4,), 4)}
            self.params = torch.arange(192, dtype=torch.float32) * fill
            self.params_lp, self.ctx, self.weights_synced = None, Ctx(), True
            self.lib = self.ctx.lib

    class Wav2Vec2ForCTC:
        def __init__(self, fill):
            self._prog = Prog(fill)
            self.variable_names = list(self._prog.info)
            self._step_seed = 3

    class Opt:
        learning_rate, beta_1, beta_2, epsilon, clipnorm = 1e-3, 0.9, 0.999, 1e-8, 1.0

        def __init__(self, it, fill):
            self.iterations = it
            self.st = {"m": torch.full((192,), 1.0 * fill), "v": torch.full((192,), 2.0 * fill)}

        def _bind(self, model):
            return self.st

    monkeypatch.setattr(RT, "stream_ptr", lambda: C.c_void_p(0))
    m1, o1 = Wav2Vec2ForCTC(1.0), Opt(2, 1.0)
    path = CK.Checkpoint(model=m1, optimizer=o1).save(str(tmp_path / "ck" / "model_step_50"))
    assert path.endswith("model_step_50-1.tsckpt")
    meta, tensors = CK.read_file(path)
    assert meta["device_salt"] == 11 and meta["step_seed"] == 3 and meta["optimizer"]["iterations"] == 2
    q = tensors["model/encoder.layers.0.attention.q_proj.kernel"]
    assert q.shape == (4, 4) and q[1, 0] == 64 + 12 and q[3, 3] == 64 + 36 + 3          # rows 12 apart inside the fused block
    assert meta["tf_keys"]["model/fe.conv0.kernel"].startswith("model/wav2vec2/feature_extractor/conv_layers/0/layer_with_weights-0/kernel")
    m2, o2 = Wav2Vec2ForCTC(0.0), Opt(0, 0.0)
    ck2 = CK.Checkpoint(model=m2, optimizer=o2)
    ck2.restore(path)
    assert ck2.save_counter == 1 and o2.iterations == 2 and m2._step_seed == 3 and calls["state_set"] == (11, 2)
    assert m2._prog.weights_synced is False
    for name, (off, shp, ld) in m1._prog.info.items():
        a = CK.gather_variables(m1._prog.params.numpy(), m1._prog.info, [name])[name]
        b = CK.gather_variables(m2._prog.params.numpy(), m2._prog.info, [name])[name]
        assert (a == b).all(), name
    assert float(m2._prog.params[76 + 8]) == 0.0                  # the v_proj columns of the fused block were not in the file: untouched
    assert float(o2.st["m"][64]) == 1.0 and float(o2.st["v"][129]) == 2.0


def test_checkpoint_file_imports_into_the_reference_model_by_tf_names(tmp_path):
    """Export -> import across the boundary: a checkpoint file holding the oracle's Wav2Vec2 weights under the product's variable
    paths (+ their TF keys, as save() records them) is assigned to the REFERENCE's own Wav2Vec2ForPreTraining (unmodified
    constructor, on the TF shim) by TF object-graph name; every reference variable must then hold exactly the tensor the
    name map of the pinning tests (oracle/ref_runner.py) associates with it. (Needs /root/reference.)"""
    import numpy as np
    import torch

    from oracle import ref_runner as R
    from oracle import wav2vec2_oracle as O

    if not R.available():
        pytest.skip("/root/reference is not present on this machine")
    ref = R.load("wav2vec2_dist")
    ocfg = O.Wav2Vec2Config("tiny")
    w = O.randomize_weights(O.init_weights(ocfg, seed=0, dtype=torch.float64), seed=1)

    class Wav2Vec2ForPreTraining:                    # stands in for the GPU model object: only the names matter for the export
        variable_names = list(w)

    tfk = CK.export_tf_names(Wav2Vec2ForPreTraining())
    path = str(tmp_path / "export-1.tsckpt")
    CK.write_file(path, {"model/" + k: v.numpy().astype(np.float32) for k, v in w.items()},
                  {"tf_keys": {"model/" + k: t for k, t in tfk.items()}})
    model = R.build_w2v(ref, "tiny", torch.randn(1, 3200, dtype=torch.float64))
    assert CK.assign_to_keras(model, path) == len(w)
    for name, var in R.w2v_variable_map(model).items():
        assert torch.equal(var.detach().to(torch.float32), w[name].to(torch.float32)), name
