set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
P="python -m pytest -q --no-header -p no:cacheprovider"
timeout 900 $P tests/test_gpu_kernels.py -m gpu > $OUT/kernels.log 2>&1; echo "kernels rc=$?"
timeout 900 $P tests/test_gpu_model.py -m gpu > $OUT/model.log 2>&1; echo "model rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_products.log 2>&1; echo "products rc=$?"
timeout 300 python bench.py --workload arxiv --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_arxiv.log 2>&1; echo "arxiv rc=$?"
timeout 300 python bench.py --workload pokec --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_pokec.log 2>&1; echo "pokec rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_products.csv python bench.py --no-cpu-baseline --no-e2e --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 20 -c 2 -f -o $OUT/gemm_nt_full python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > $OUT/ncu_gemm_nt.log 2>&1; echo "gemm_nt full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd -s 6 -c 2 -f -o $OUT/bn_bwd_full python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 3 > $OUT/ncu_bn_bwd.log 2>&1; echo "bn_bwd full rc=$?"
grep -E "passed|failed" $OUT/kernels.log $OUT/model.log
grep -E "^(FAILED|E   [A-Za-z])" $OUT/kernels.log | head -30
grep -E "^(FAILED|E   [A-Za-z])" $OUT/model.log | head -30
for f in products arxiv pokec; do grep "^{" $OUT/bench_$f.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline'] or {}
    print('$f', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'spmm frac', r.get('frac'), 'share', r.get('share_of_step'), 'e2e ms', (d['e2e'] or {}).get('ms_per_step'))"; tail -n 3 $OUT/bench_$f.log | grep -v "^{" | cut -c1-300; done
