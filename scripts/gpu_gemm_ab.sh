# gemm_nt A/B on one B200: kernel tests, per-launch timing of the step's GEMM shapes (old r1b kernel vs both schedules of the
# current one), then the products-shaped step
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" > $OUT/gemm_tests.log 2>&1; echo "gemm tests rc=$?"
grep -E "passed|failed" $OUT/gemm_tests.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/gemm_tests.log | cut -c1-300 | head -20
[ -f sgformer_b200/lib/libsgformer_b200_r1b.so ] && timeout 300 python scripts/bench_gemm_nt.py --lib sgformer_b200/lib/libsgformer_b200_r1b.so --schedules 0 2>&1 | tee $OUT/gemm_micro_r1b.log
timeout 300 python scripts/bench_gemm_nt.py --schedules 1,2 2>&1 | tee $OUT/gemm_micro.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > $OUT/bench_auto.log 2>&1; echo "bench rc=$?"
grep "^{" $OUT/bench_auto.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'spmm', d['roofline']['avg_launch_ms'])"
