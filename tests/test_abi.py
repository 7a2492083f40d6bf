"""CPU-side checks of the drop-in boundary: libtethys.so loads without a GPU or libcuda, exports every symbol
include/tethys.h declares, the ctypes table covers them all, and the product path fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tethys.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ts_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert "ts_gemm" in syms and "ts_w2v_forward" in syms and "ts_optim_step" in syms
    assert len(syms) >= 20


def test_library_loads_and_exports_every_declared_symbol():
    from tethys_speech_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "libtethys.so missing: run make / __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/tethys.h but not exported: {missing}"
    assert lib.ts_version() == 100


def test_ctypes_table_matches_header():
    from tethys_speech_b200 import _lib

    declared = set(_declared_symbols())
    bound = set(_lib.SYMBOLS)
    assert declared == bound, f"only in header: {declared - bound}; only in ctypes table: {bound - declared}"
    _lib.load()


def _c_layout(struct, fields):
    """sizeof / offsetof of a struct of include/tethys.h as gcc lays it out (a tiny C program compiled on the spot)."""
    import subprocess
    import tempfile

    src = '#include <stdio.h>\n#include <stddef.h>\n#include "tethys.h"\nint main(void) {\n'
    src += f'  printf("%zu\\n", sizeof({struct}));\n'
    for f in fields:
        src += f'  printf("%zu\\n", offsetof({struct}, {f}));\n'
    src += "  return 0;\n}\n"
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "l.c"), os.path.join(d, "l")
        open(c, "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    return int(out[0]), [int(x) for x in out[1:]]


@pytest.mark.parametrize("cname,mirror", [("ts_gemm_desc", "GemmDesc"), ("ts_attn_desc", "AttnDesc"), ("ts_w2v_config", "W2VConfig"),
                                          ("ts_whisper_config", "WhisperCfg"), ("ts_step_args", "StepArgs")])
def test_struct_layouts_match_c(cname, mirror):
    """Every ctypes mirror in _lib.py has the size and the field offsets gcc gives the struct of include/tethys.h."""
    from tethys_speech_b200 import _lib

    m = getattr(_lib, mirror)
    names = [f[0] for f in m._fields_]
    size, offs = _c_layout(cname, names)
    assert ctypes.sizeof(m) == size, (cname, ctypes.sizeof(m), size)
    assert [getattr(m, n).offset for n in names] == offs, cname


def test_integration_md_struct_snippet_is_current():
    """The ctypes struct a maintainer would copy out of INTEGRATION.md is, field for field, the one the library reads."""
    from tethys_speech_b200 import _lib

    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class ts_w2v_config\(C\.Structure\):.*?\n(    _fields_ = \[.*?\])\n", txt, flags=re.S)
    assert m, "INTEGRATION.md no longer shows the ts_w2v_config binding"
    ns = {"C": ctypes}
    exec("class S(C.Structure):\n" + m.group(1), ns)
    doc = ns["S"]
    assert [f[0] for f in doc._fields_] == [f[0] for f in _lib.W2VConfig._fields_]
    assert ctypes.sizeof(doc) == ctypes.sizeof(_lib.W2VConfig)
    assert [getattr(doc, f[0]).offset for f in doc._fields_] == [getattr(_lib.W2VConfig, f[0]).offset for f in _lib.W2VConfig._fields_]


def test_integration_md_step_args_snippet_is_current():
    """Same check for the ts_step_args binding INTEGRATION.md shows for the composite step entries."""
    from tethys_speech_b200 import _lib

    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class ts_step_args\(C\.Structure\):.*?\n(    _fields_ = \[.*?\])\n", txt, flags=re.S)
    assert m, "INTEGRATION.md no longer shows the ts_step_args binding"
    ns = {"C": ctypes}
    exec("class S(C.Structure):\n" + m.group(1), ns)
    doc = ns["S"]
    assert [f[0] for f in doc._fields_] == [f[0] for f in _lib.StepArgs._fields_]
    assert [getattr(doc, f[0]).offset for f in doc._fields_] == [getattr(_lib.StepArgs, f[0]).offset for f in _lib.StepArgs._fields_]
    assert ctypes.sizeof(doc) == ctypes.sizeof(_lib.StepArgs)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tethys_speech_b200 import _lib

    with pytest.raises(_lib.TethysError):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tethys_speech_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt:
                    offenders.append(f)
    assert not offenders, f"product files referencing oracle/: {offenders}"


def test_nvml_sampler_side_car_builds_and_fails_cleanly_without_a_driver():
    """the reference's only native component (NVML/NVML.cpp) re-provided as tools/nvml_sampler. Without a
    driver it must exit 2 with a message (NVML is resolved at run time); with one it prints one line per GPU."""
    import subprocess

    exe = os.path.join(ROOT, "tools", "nvml_sampler")
    if not os.path.exists(exe):
        pytest.skip("tools/nvml_sampler not built (run make)")
    r = subprocess.run([exe, "--count", "1"], capture_output=True, text=True, timeout=30)
    assert r.returncode in (0, 1, 2), r
    if r.returncode == 0:
        assert "GPU Util:" in r.stdout
    else:
        assert "nvml_sampler" in r.stderr
