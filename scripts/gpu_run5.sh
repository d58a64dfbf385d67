set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
P="python -m pytest -q --no-header -p no:cacheprovider"
timeout 900 $P tests -m gpu > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
for wl in products papers100M-minibatch pokec; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$wl.log 2>&1; echo "$wl rc=$?"
done
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -20
for f in products papers100M-minibatch pokec; do grep "^{" $OUT/bench_$f.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline'] or {}
    print('$f', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'graph', d['config'].get('cuda_graph'), 'spmm frac', round(r.get('frac',0),3), 'share', round(r.get('share_of_step',0),3))"; grep -E "capture failed|Error" $OUT/bench_$f.log | head -3; done
