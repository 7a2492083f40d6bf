set -x
cd $GRAFT_REPO_ROOT
timeout -k 5 900 python -m pytest tests -q -m gpu -s > gpurun_out/r02_pytest_gpu_b.log 2>&1; echo "pytest rc=$?"
grep -E "^\[|passed|failed|^E  " gpurun_out/r02_pytest_gpu_b.log | cut -c1-700 | tail -80
timeout 900 python tools/parity_report.py --out gpurun_out/r02_parity_report_b.json --cases w2v_tiny_bf16,w2v_tiny_bf16_seed3,w2v_tiny_bf16_b4_1s,w2v_tiny_bf16_T200,w2v_small_bf16_2s,w2v_base_bf16_2s,w2v_base_bf16_15s_b1,whisper_smallcfg_bf16,whisper_tiny_bf16,whisper_default_bf16_30s_b1 > gpurun_out/r02_parity_b.log 2>&1; echo "parity rc=$?"
