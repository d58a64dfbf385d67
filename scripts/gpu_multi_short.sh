# 8-GPU refresh of the headline multi-GPU numbers: dp (weak) products, mini-batch dp, row-sharded products
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_dp_$N.log 2>&1; echo "dp rc=$?"
timeout 300 $T bench.py --gpus $N --steps 5 --warmup 3 --workload papers100M-minibatch > $OUT/bench_mb_$N.log 2>&1; echo "mb rc=$?"
timeout 300 $T bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --parallel rows > $OUT/bench_rows_$N.log 2>&1; echo "rows products rc=$?"
for f in bench_dp_$N bench_mb_$N bench_rows_$N; do grep "^{" $OUT/$f.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$f', {k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')})"; grep -E "Error|error" $OUT/$f.log | head -3; done
