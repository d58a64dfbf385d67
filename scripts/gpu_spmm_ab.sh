# SpMM tuning variants (sgformer_b200/_build.build_variant) on the products-shaped graph, 1 GPU
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "spmm" > $OUT/spmm_tests.log 2>&1; echo "spmm tests rc=$?"
grep -E "passed|failed" $OUT/spmm_tests.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/spmm_tests.log | cut -c1-300 | head
timeout 200 python scripts/bench_spmm.py 2>&1 | tail -1 | tee $OUT/spmm_variants.log
for l in sgformer_b200/lib/libsgformer_b200_spmm_*.so; do timeout 200 python scripts/bench_spmm.py --lib $l 2>&1 | tail -1 | tee -a $OUT/spmm_variants.log; done
