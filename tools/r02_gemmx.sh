cd $GRAFT_REPO_ROOT
for env in "A=0" "TETHYS_GEMM_RASTER=1" "TETHYS_GEMM_BN=128" "TETHYS_GEMM_BN=256" "TETHYS_GEMM_BN=64"; do
echo "=== $env"
env $env timeout 200 ./tools/selftest_gemm time 2>&1 | grep "time" | grep -v "8192\|conv" | cut -c1-170
done
