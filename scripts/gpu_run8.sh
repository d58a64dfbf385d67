set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
P="python -m pytest -q --no-header -p no:cacheprovider --durations=8"
timeout 900 $P tests -m gpu > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_products.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --rmat > $OUT/bench_products_rmat.log 2>&1
SGF_FUSED_STATS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_products_fused.log 2>&1
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -20; grep -A9 "slowest" $OUT/all_gpu.log | cut -c1-120
for f in products products_rmat products_fused; do grep "^{" $OUT/bench_$f.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline'] or {}
    print('$f', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'nnz', d['config'].get('nnz_per_gpu'), 'spmm frac', round(r.get('frac',0),3), 'ms', round(r.get('avg_launch_ms',0),3), 'share', round(r.get('share_of_step',0),3))"; grep -E "capture failed|Error" $OUT/bench_$f.log | head -3; done
