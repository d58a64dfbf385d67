# r2m (1 GPU): BatchNorm backward with the cp.async staging ring: whole GPU suite with the ring on, bench A/B
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
SGF_BN_BWD_RING=1 timeout 600 python -m pytest tests -q -m gpu -x > $OUT/r2m_tests.log 2>&1; echo "pytest (ring on) rc=$?"
grep -E "passed|failed" $OUT/r2m_tests.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/r2m_tests.log | cut -c1-300 | head -20
for R in 0 1; do
  SGF_BN_BWD_RING=$R timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_ring$R.log 2>&1
  grep "^{" $OUT/bench_ring$R.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('SGF_BN_BWD_RING=$R ms/step', d['ms_per_step'], 'spmm', d['roofline']['avg_launch_ms'], 'clk', d['clocks']['sm_mhz'])"
done
