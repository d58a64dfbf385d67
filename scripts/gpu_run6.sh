set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
P="python -m pytest -q --no-header -p no:cacheprovider"
timeout 900 $P tests -m gpu > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
timeout 600 python bench.py --workload papers100M-minibatch --steps 10 --warmup 3 > $OUT/bench_mb.log 2>&1; echo "mb rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_products.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_products.csv python bench.py --no-cpu-baseline --no-e2e --no-graph --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launches rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -20
grep "^{" $OUT/bench_mb.log | cut -c1-700; grep -E "Error" $OUT/bench_mb.log | head
grep "^{" $OUT/bench_products.log | cut -c1-200
