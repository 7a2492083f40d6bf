cd $GRAFT_REPO_ROOT
for args in "6000 512 512 0 0 1" "6000 512 512 0 0 1 cold" "6000 512 2048 0 0 1" "6000 1536 512 0 1 1" "400 512 512 0 0 1" "400 512 2048 0 0 1" "6000 768 768 0 0 1"; do
timeout 120 python tools/gemm_trace.py $args 2>&1 | grep -v "^\s*\.\.\." | head -9
done
