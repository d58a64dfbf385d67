# gemm_nt A/B on one B200: per-launch timing of the step's GEMM shapes (AUTO / streaming / resident-B), all GPU tests, the
# products-shaped step and its launch list
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/bench_gemm_nt.py --schedules 0,1,2 2>&1 | tee $OUT/gemm_micro.log
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -20
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > $OUT/bench_auto.log 2>&1; echo "bench rc=$?"
grep "^{" $OUT/bench_auto.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'spmm', d['roofline']['avg_launch_ms'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_products.csv python bench.py --no-cpu-baseline --no-e2e --no-graph --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launches rc=$?"
