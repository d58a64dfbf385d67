# r2d (1 GPU): final check of the tree: GPU tests, smoke, default bench line (resident single-block attention apply on / off)
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu --durations=3 > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 $OUT/smoke.log
for R in 0 1; do
  SGF_ATTN_RESIDENT=$R timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_attnres$R.log 2>&1
  grep "^{" $OUT/bench_attnres$R.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('SGF_ATTN_RESIDENT=$R ms/step', d['ms_per_step'], 'spmm', d['roofline']['avg_launch_ms'], 'clk', d['clocks'])"
done
for W in 0 32 48 64 96; do SGF_CSR_FILL_WINDOW_MB=$W timeout 200 python scripts/bench_csr.py 2>&1 | tail -n 1; done
for P in 0 1; do
  SGF_BENCH_PREPARE=$P timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $OUT/bench_prepare$P.log 2>&1
  grep "^{" $OUT/bench_prepare$P.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('SGF_BENCH_PREPARE=$P ms/step', d['ms_per_step'], 'e2e', d['e2e'])"
done
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_default.log 2>&1; echo "bench rc=$?"
grep "^{" $OUT/bench_default.log | cut -c1-6000
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_nt_kernel -s 45 -c 15 --csv --log-file $OUT/r2d_gemm_nt_times.csv python bench.py --no-cpu-baseline --no-e2e --no-extra --no-graph --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu gemm_nt times rc=$?"
