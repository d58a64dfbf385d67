# r2b2 (N GPUs): NCCL row-sharding parity (pytest wrapper), push vs all-gather for the row-sharded products / Pokec steps
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > $OUT/multi_pytest_$N.log 2>&1; echo "pytest multi rc=$?"; grep -E "passed|failed|skipped" $OUT/multi_pytest_$N.log
grep -E "^\[(fp32|bf16)\]|multi_gpu_check" $OUT/multi_pytest_$N.log | head; grep -E "Error|error" $OUT/multi_pytest_$N.log | head -5 | cut -c1-300
for MODE in ${MODES:-push allgather}; do
  for WL in ${WLS:-products pokec}; do
    SGF_C4_MODE=$MODE timeout 500 $T bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --parallel rows --workload $WL > $OUT/bench_rows_${WL}_${N}_$MODE.log 2>&1; echo "rows $WL $MODE rc=$?"
    grep "^{" $OUT/bench_rows_${WL}_${N}_$MODE.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$WL $MODE', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'graph', d['config']['cuda_graph'], 'spmm ms', d['roofline'] and d['roofline']['avg_launch_ms'])"
    grep -E "Error|error|Warning: sgformer|capture failed" $OUT/bench_rows_${WL}_${N}_$MODE.log | head -5 | cut -c1-300
  done
done
SGF_BENCH_MULTI_GRAPH=0 SGF_C4_MODE=push timeout 500 $T bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --parallel rows --workload products > $OUT/bench_rows_products_${N}_push_eager.log 2>&1; echo "rows products push eager rc=$?"
grep "^{" $OUT/bench_rows_products_${N}_push_eager.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('products push eager', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'spmm ms', d['roofline'] and d['roofline']['avg_launch_ms'])"
