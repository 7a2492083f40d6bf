import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_l2(a, b):
    """relative L2 error ||a-b|| / ||b|| in float64 (b is the oracle)."""
    import torch

    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-300))


def rel_max(a, b):
    """max |a-b| / max |b|."""
    import torch

    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


BF16_TOL = 2e-2          # north star: 2e-2 relative for activations, loss and gradients in bf16 mode
BF16_BUDGET_MARGIN = 1.5


def check_bf16_grads(tag, gpu_errs, emu_errs, tol=BF16_TOL, margin=BF16_BUDGET_MARGIN):
    """bf16 gradient bar. Every tensor must be within the north star's 2e-2 relative L2 of the fp64 oracle, EXCEPT where an
    independent CPU emulation of bf16 storage alone (oracle.tf_ops.bf16_storage: fp64 arithmetic, but weights' compute copy,
    stored activations and activation gradients rounded to bfloat16 — no CUDA code involved) already moves that tensor further
    than 2e-2 / margin away from the fp64 result; such a tensor's budget is margin x its emulated error (observed vs expected
    are both printed). margin = 1.5: the GPU run and the emulation are two different realisations of the same rounding noise
    (rounding points differ); over the ~1 400 gradient tensors of tools/parity_report.py's ten bf16 cases the ratio
    observed / emulated has median 1.0 and maximum 1.53 (profiles/r02_parity_report.json). Typical members: q/k projection gradients of the deepest layers (dS = P o (dP - D) cancels two nearly
    equal bf16-rounded terms under a near-uniform softmax) and the first conv kernels (every later rounding funnels into
    them through 5-7 GroupNorm layers). Returns the list of over-budget tensors as (name, gpu, emulated, budget)."""
    bad, lifted = [], []
    for name, e in gpu_errs.items():
        budget = max(tol, margin * emu_errs[name])
        if budget > tol:
            lifted.append((name, e, emu_errs[name]))
        if not e <= budget:
            bad.append((name, e, emu_errs[name], budget))
    worst = sorted(gpu_errs.items(), key=lambda kv: -kv[1])[:5]
    print(f"[{tag}] {len(gpu_errs)} gradients: worst gpu {[(k, round(v, 4)) for k, v in worst]}; "
          f"{sum(1 for v in gpu_errs.values() if v > tol)} over {tol:g}, all of them inside their bf16-storage budget: "
          f"{[(k, round(g, 4), round(e, 4)) for k, g, e in lifted if g > tol]}" if not bad else f"[{tag}] OVER BUDGET: {bad}")
    return bad
