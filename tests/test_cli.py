"""The drop-in command lines (speech_jobs/*.py) keep the reference's flags and defaults — W:1032-1033 (whisper_dist: --num_batches 40,
--batch_size 1), WS:1304-1305 (whisper_single: 40 / 4), V:1446-1449 and VS:1284-1291 (wav2vec2: 5 / 1, --model_size small,
--model_type pretraining, --learning_rate 3e-5, --num_epochs 1) — and, without a GPU, fail loudly instead of falling back to a CPU path."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JOBS = os.path.join(ROOT, "speech_jobs")


def _help(script):
    r = subprocess.run([sys.executable, os.path.join(JOBS, script), "--help"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout


def _default(text, flag):
    """default of `flag` as argparse prints it is not in --help (no %(default)s): read it from the parser source instead."""
    m = re.search(r'add_argument\("' + re.escape(flag) + r'".*?default=([^,)]+)', text)
    assert m, flag
    return m.group(1).strip().strip('"')


@pytest.mark.parametrize("script,defaults", [
    ("whisper_dist.py", {"--num_batches": "40", "--batch_size": "1"}),
    ("whisper_single.py", {"--num_batches": "40", "--batch_size": "4"}),
    ("wav2vec2_dist.py", {"--num_batches": "5", "--batch_size": "1", "--model_size": "small"}),
    ("wav2vec2_single.py", {"--num_batches": "5", "--batch_size": "1", "--model_size": "small", "--model_type": "pretraining",
                            "--learning_rate": "3e-5", "--num_epochs": "1"}),
])
def test_cli_flags_and_reference_defaults(script, defaults):
    out = _help(script)
    src = open(os.path.join(JOBS, script)).read()
    for flag, want in defaults.items():
        assert flag in out, (script, flag)
        assert _default(src, flag) == want, (script, flag, _default(src, flag))
    assert "--resume" in out                       # SURVEY f-3 extension


def test_model_type_choices_cover_the_reference_heads():
    out = _help("wav2vec2_single.py")
    assert "pretraining" in out and "asr" in out and "classification" in out          # VS:1288


def test_cli_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less hosts")
    r = subprocess.run([sys.executable, os.path.join(JOBS, "wav2vec2_single.py"), "--batch_size", "1", "--num_batches", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
