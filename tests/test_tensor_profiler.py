"""the Tiresias tensor-size profiler (tethys_speech_b200/tensor_profiler.py) — same bookkeeping, files and numbers
as the reference's TensorProfiler (WT:20-458). Host logic: runs without a GPU."""
import json
import os

import numpy as np
import pytest


def test_profiler_bookkeeping_files_and_tiresias_number(tmp_path):
    from tethys_speech_b200.tensor_profiler import TensorProfiler

    p = TensorProfiler(log_dir=str(tmp_path), verbose=False)
    sizes = []
    for step in range(8):
        p.start_step(step)
        p.log_tensor_size(np.zeros((4, 80, 3000), np.float32), "input_0", "input")
        p.log_tensor_size(np.zeros((4, 1500, 768 + step), np.float32), "encoder_out", "activation")     # grows: steps differ
        p.log_tensor_size(None, "nothing")
        p.log_gradients([np.zeros((768, 768), np.float32), None], ["a/kernel:0", "b/kernel:0"])
        p.log_memory_usage()
        sizes.append(p.end_step())
    want = [(4 * 80 * 3000 + 4 * 1500 * (768 + s) + 768 * 768) * 4 / 2 ** 20 for s in range(8)]
    assert np.allclose(sizes, want)
    assert abs(p.get_tiresias_tensorsize() - np.mean(want[2:])) < 1e-9         # warm-up = min(3, 8 // 4) = 2 steps (WT:213)
    s = p.save_final_results()
    p.close()
    assert s["total_steps"] == 8 and s["total_operations"] == 8 * 3
    assert s["operation_stats"]["gradient_a/kernel:0"]["count"] == 8
    from scipy import stats

    all_mb = [d["size_mb"] for d in p.tensor_details]
    assert abs(s["model_skewness"] - float(stats.skew(all_mb))) < 1e-12
    assert set(s["skewness_analysis"]["layer_type_skewness"]) == {"input", "activation", "gradient"}
    for f in ("tensor_sizes.txt", "memory_usage.txt", "summary.txt", "tiresias_tensorsize.txt", "final_summary.json", "tiresias_result.json",
              "legacy_skewness_result.txt"):
        assert os.path.getsize(tmp_path / f) > 0
    assert open(tmp_path / "tensor_sizes.txt").readline().strip() == "step,operation,tensor_type,size_bytes,size_mb,shape"
    assert open(tmp_path / "tiresias_tensorsize.txt").read().splitlines()[1] == f"0,{want[0]:.4f}"
    r = json.load(open(tmp_path / "tiresias_result.json"))
    assert r["measurement_method"] == "Tiresias_style" and abs(r["tensorsize_mb"] - np.mean(want[2:])) < 1e-9


@pytest.mark.gpu
def test_profile_step_on_the_whisper_step(tmp_path):
    import torch

    from tethys_speech_b200 import whisper as W
    from tethys_speech_b200.tensor_profiler import TensorProfiler, profile_step

    cfg = W.WhisperConfig()
    cfg.d_model, cfg.d_ff, cfg.encoder_layers, cfg.decoder_layers = 128, 256, 2, 2
    cfg.encoder_attention_heads = cfg.decoder_attention_heads = 2
    cfg.vocab_size, cfg.n_ctx, cfg.decoder_start_token_id = 512, 64, 500
    model = W.WhisperForConditionalGeneration(cfg, precision="bf16", seed=0)
    opt = W.Adam(learning_rate=1e-4)
    f = torch.randn(2, 80, 128)
    lab = torch.randint(0, 100, (2, 20), dtype=torch.int32)
    p = TensorProfiler(log_dir=str(tmp_path), verbose=False)
    for step in range(4):
        loss, mb = profile_step(p, step, model, (f, lab), lambda: W.train_step(model, (f, lab), opt))
    s = p.save_final_results()
    p.close()
    n_param_bytes = sum(v.numel() * 4 for v in model.trainable_variables)
    assert mb * 2 ** 20 >= 2 * n_param_bytes                      # gradients + parameters are in every step's total
    assert s["total_steps"] == 4 and "logits" in s["operation_stats"] and np.isfinite(float(loss))
