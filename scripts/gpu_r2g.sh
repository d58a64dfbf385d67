# r2g (1 GPU): ncu --set full of the graph-build kernels (bucketed and direct)
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
SGF_CSR_BUCKETS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bucket_scatter|bucket_build|bucket_count" -s 6 -c 3 -o $OUT/r2g_bucket python scripts/bench_csr.py > /dev/null 2>&1; echo "ncu bucket rc=$?"
SGF_CSR_BUCKETS=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"csr_fill|csr_count|csr_sort_rows_warp" -s 36 -c 6 -o $OUT/r2g_direct python scripts/bench_csr.py > /dev/null 2>&1; echo "ncu direct rc=$?"
ls -la $OUT/r2g_*
