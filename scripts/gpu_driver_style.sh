# what the driver runs at N GPUs: one torchrun of bench.py (headline dp line + the `extra` block measured in child processes)
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "${SKIP_MULTI:-0}" != "1" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > $OUT/multi_pytest_$N.log 2>&1; echo "pytest multi rc=$?"; grep -E "passed|failed|skipped" $OUT/multi_pytest_$N.log; fi
timeout 1500 $T bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 ${BENCH_FLAGS:-} > $OUT/bench_driver_$N.log 2> $OUT/bench_driver_$N.err; echo "bench --gpus $N rc=$?"
grep "^{" $OUT/bench_driver_$N.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    print('headline', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'graph', d['config']['cuda_graph'])
    for k,v in (d.get('extra') or {}).items(): print('  extra', k, json.dumps(v)[:420])"
grep -E "Elapsed|Error|error" $OUT/bench_driver_$N.err | head -8 | cut -c1-200
for spec in ${AB:-}; do
  WL=${spec%%:*}; MODE=${spec##*:}
  SGF_C4_MODE=$MODE timeout 500 $T bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --parallel rows --workload $WL > $OUT/bench_rows_${WL}_${N}_$MODE.log 2>&1; echo "rows $WL $MODE rc=$?"
  grep "^{" $OUT/bench_rows_${WL}_${N}_$MODE.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$WL $MODE', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'graph', d['config']['cuda_graph'])"
done
