# row-sharded (C1-C5) mode on N GPUs: parity check against the single-GPU run, then the products-shaped step with the C4
# all-gather blocking (SGF_C4_CHUNKS=1) and pipelined in 2 / 4 column chunks, then the Pokec-shaped step
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tests/multi_gpu_check.py > $OUT/multi_check_$N.log 2>&1; echo "multi_check rc=$?"
tail -n 4 $OUT/multi_check_$N.log
for c in ${CHUNKS:-1 2 4}; do
  SGF_C4_CHUNKS=$c timeout 400 $T bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --parallel rows > $OUT/bench_rows_${N}_c$c.log 2>&1; echo "rows products chunks=$c rc=$?"
  grep "^{" $OUT/bench_rows_${N}_c$c.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('chunks=$c', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')})"; grep -E "Error|error" $OUT/bench_rows_${N}_c$c.log | head -3
done
timeout 400 $T bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --parallel rows --workload pokec > $OUT/bench_rows_pokec_$N.log 2>&1; echo "rows pokec rc=$?"
grep "^{" $OUT/bench_rows_pokec_$N.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('pokec', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')})"
