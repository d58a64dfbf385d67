set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -20
for WL in arxiv; do
  timeout 300 python bench.py --workload $WL --no-cpu-baseline --no-e2e --steps 20 --warmup 5 > $OUT/bench_$WL.log 2>&1; echo "bench $WL rc=$?"
  grep "^{" $OUT/bench_$WL.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$WL ms_per_step', d['ms_per_step'], 'nodes/s', d['value'], 'launches/step', d['gpu_launches']/d['steps'])"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_$WL.csv python bench.py --workload $WL --no-cpu-baseline --no-e2e --no-graph --steps 2 --warmup 3 > $OUT/ncu_launches_$WL.log 2>&1; echo "launches $WL rc=$?"
done
