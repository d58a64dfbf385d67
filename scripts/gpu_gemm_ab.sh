# A/B of the gemm_nt schedules on the products-shaped step (1 GPU): auto (resident B when it fits), streaming, forced resident
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu --durations=5 > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -20
for s in 0 1 2; do
  SGF_GEMM_NT_SCHEDULE=$s timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > $OUT/bench_sched$s.log 2>&1; echo "bench sched=$s rc=$?"
  grep "^{" $OUT/bench_sched$s.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['ms_per_step'], 'spmm', d['roofline']['avg_launch_ms'])"
done
for s in 0 2; do
SGF_GEMM_NT_SCHEDULE=$s timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_products_sched$s.csv python bench.py --no-cpu-baseline --no-e2e --no-graph --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launches rc=$?"
done
