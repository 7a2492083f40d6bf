cd $GRAFT_REPO_ROOT
for args in "6000 768 768 0 1 1" "6000 768 768 0 1 1 cold" "6000 768 768 0 1 2 cold" "6000 3072 768 0 1 1 cold" "6000 3072 768 0 1 2 cold" "6000 768 3072 0 1 1 cold" "6000 768 3072 0 1 2 cold" "8192 8192 8192 0 0 2 cold"; do
  timeout 120 python tools/gemm_trace.py $args 2>&1 | tail -12
done > gpurun_out/r02_gemm_trace.log 2>&1
cat gpurun_out/r02_gemm_trace.log
