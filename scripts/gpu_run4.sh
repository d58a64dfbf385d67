set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
P="python -m pytest -q --no-header -p no:cacheprovider"
timeout 900 $P tests -m gpu -x > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
for wl in products arxiv pokec papers-batch tiny; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$wl.log 2>&1; echo "$wl rc=$?"
done
timeout 600 python bench.py --workload arxiv --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/bench_arxiv_nograph.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_products.csv python bench.py --no-cpu-baseline --no-e2e --no-graph --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launches rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -20
tail -n 2 $OUT/smoke.log
for f in products arxiv pokec papers-batch tiny arxiv_nograph; do grep "^{" $OUT/bench_$f.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline'] or {}
    print('$f', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'graph', d['config'].get('cuda_graph'), 'spmm frac', round(r.get('frac',0),3), 'share', round(r.get('share_of_step',0),3), 'e2e ms', (d['e2e'] or {}).get('ms_per_step'))"; grep -E "capture failed|Error" $OUT/bench_$f.log | head -3; done
