# r2e (1 GPU): graph-build speedups (window sweep), prefetched graph build in the e2e loop (A/B), final default bench line
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q -m gpu -k "feeder or csr or symmetry or dropout" > $OUT/r2e_tests.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" $OUT/r2e_tests.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/r2e_tests.log | cut -c1-300 | head -20
for W in 0 32 48 64 96; do SGF_CSR_FILL_WINDOW_MB=$W timeout 200 python scripts/bench_csr.py 2>&1 | tail -n 1; done
for P in 0 1; do
  SGF_BENCH_PREPARE=$P timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $OUT/bench_prepare$P.log 2>&1
  grep "^{" $OUT/bench_prepare$P.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('SGF_BENCH_PREPARE=$P ms/step', d['ms_per_step'], 'e2e', d['e2e'])"
done
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_default.log 2>&1; echo "bench rc=$?"
grep "^{" $OUT/bench_default.log | cut -c1-1500
