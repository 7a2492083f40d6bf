#!/usr/bin/env bash
# compute-sanitizer passes over the hand-written kernels on SMALL shapes (the reference has no such check).
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck]        (default: memcheck)       -> gpurun_out/sanitize_<tool>.log
# Runs the two standalone self-tests (tcgen05 GEMM engine incl. CTA-pair tiles and split-K; fused attention fwd/bwd) and the
# tiny-preset parity tests of both train steps under the sanitizer. Needs a B200; expect ~20-50x slowdown (a few minutes).
# NOT run in round 2 (the round's GPU budget was spent on measurement); kept so that the next change to a kernel can be checked
# with one command:  gpurun --timeout 900 -- 'make -j && tools/sanitize.sh memcheck'
set -u
TOOL="${1:-memcheck}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/gpurun_out/sanitize_${TOOL}.log"
mkdir -p "$ROOT/gpurun_out"
: > "$OUT"
SAN="compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20"
rc=0
for cmd in "$ROOT/tools/selftest_gemm quick" "$ROOT/tools/selftest_attn quick"; do
  echo "== $SAN $cmd" | tee -a "$OUT"
  timeout 600 $SAN $cmd >> "$OUT" 2>&1 || rc=$?
done
echo "== $SAN python -m pytest (tiny presets)" | tee -a "$OUT"
( cd "$ROOT" && TETHYS_PDL=0 timeout 1500 $SAN --target-processes all python -m pytest -q -m gpu -x \
    "tests/test_w2v_gpu.py::test_w2v_tiny_fp32_forward_backward" "tests/test_w2v_gpu.py::test_w2v_tiny_bf16_forward_backward" \
    "tests/test_whisper_gpu.py::test_whisper_small_config_fp32" "tests/test_whisper_gpu.py::test_whisper_small_config_bf16" ) >> "$OUT" 2>&1 || rc=$?
grep -E "ERROR SUMMARY|passed|failed" "$OUT" | tail -8
exit $rc
