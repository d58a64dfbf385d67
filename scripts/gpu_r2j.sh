# r2j (1 GPU): final state of the tree: all GPU tests, smoke, graph-build timing, default bench line, ncu launch list of the same command
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu --durations=3 > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 $OUT/smoke.log
for W in 96 128; do SGF_CSR_FILL_WINDOW_MB=$W timeout 200 python scripts/bench_csr.py 2>&1 | tail -n 1; done
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_default.log 2>&1; echo "bench rc=$?"
grep "^{" $OUT/bench_default.log | cut -c1-7000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2j_launches_products.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-graph > $OUT/r2j_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
