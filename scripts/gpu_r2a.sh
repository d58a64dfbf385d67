# r2a: first GPU check of the Gram-form attention path: GPU tests, products-shaped bench with and without it, launch list
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu --durations=5 > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 $OUT/smoke.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_gram.log 2>&1; echo "bench gram rc=$?"
grep "^{" $OUT/bench_gram.log | cut -c1-3000
SGF_GRAM_ATTENTION=0 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_legacy.log 2>&1; echo "bench legacy rc=$?"
grep "^{" $OUT/bench_legacy.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/r2a_launches_products.csv python bench.py --no-cpu-baseline --no-e2e --no-extra --no-graph --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launches rc=$?"
tail -n 3 $OUT/ncu_launches.log | cut -c1-300
