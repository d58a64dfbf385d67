# r2b (N GPUs, default 2): GPU test suite (incl. the NCCL row-sharding parity test), then the row-sharded products / Pokec steps
# with the pushed halo exchange (copy-engine peer copies + flagged SpMM) against the blocking all-gather
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -30
timeout 300 $T tests/multi_gpu_check.py > $OUT/multi_check_$N.log 2>&1; echo "multi_check rc=$?"; tail -n 5 $OUT/multi_check_$N.log | cut -c1-300
for MODE in push allgather; do
  for WL in products pokec; do
    SGF_C4_MODE=$MODE timeout 500 $T bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --parallel rows --workload $WL > $OUT/bench_rows_${WL}_${N}_$MODE.log 2>&1; echo "rows $WL $MODE rc=$?"
    grep "^{" $OUT/bench_rows_${WL}_${N}_$MODE.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$WL $MODE', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'spmm ms', d['roofline'] and d['roofline']['avg_launch_ms'])"
    grep -E "Error|error|Warning: sgformer" $OUT/bench_rows_${WL}_${N}_$MODE.log | head -5 | cut -c1-300
  done
done
for B in 2 3; do
  SGF_LNATTN_BLOCKS=$B timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_lnattn$B.log 2>&1
  grep "^{" $OUT/bench_lnattn$B.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('single GPU, ln_bwd_attn CTAs/SM=$B', d['ms_per_step'])"
done
