# Round-end verification on one B200 (what the driver runs, plus the numbers quoted in DESIGN.md / profiles):
#   GPU tests, smoke(), default bench (with cpu_baseline + e2e), the reference arm, launch list of the products step,
#   the other BASELINE shapes, preprocessing / evaluation kernels at the products shape.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu --durations=5 > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_default.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_reference.log 2>&1; echo "reference rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_products.csv python bench.py --no-cpu-baseline --no-e2e --no-graph --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launches rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -20; grep -A6 "slowest" $OUT/all_gpu.log | cut -c1-120
tail -n 2 $OUT/smoke.log
grep "^{" $OUT/bench_default.log
grep "^{" $OUT/bench_reference.log | cut -c1-600
for WL in arxiv pokec papers100M-minibatch; do
  timeout 300 python bench.py --workload $WL --no-cpu-baseline --no-e2e --steps 20 --warmup 5 > $OUT/bench_$WL.log 2>&1; echo "bench $WL rc=$?"
  grep "^{" $OUT/bench_$WL.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$WL ms_per_step', d['ms_per_step'], 'nodes/s', d['value'])"
done
timeout 300 python bench.py --rmat --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > $OUT/bench_products_rmat.log 2>&1; echo "bench rmat rc=$?"
grep "^{" $OUT/bench_products_rmat.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rmat ms_per_step', d['ms_per_step'], 'spmm', d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
timeout 600 python scripts/bench_prep.py 2>&1 | tee $OUT/prep_products.log
