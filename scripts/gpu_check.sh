#!/bin/bash
# Runs on the GPU box (via gpurun): kernel tests in isolated processes, model tests, smoke, short bench.
# Logs go to gpurun_out/; a summary is printed at the end (gpurun only shows the tail).
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
run() { # name timeout cmd...
  local name=$1; shift; local to=$1; shift
  timeout $to "$@" > $OUT/$name.log 2>&1
  echo "$name rc=$?" >> $OUT/summary.txt
}
: > $OUT/summary.txt
P="python -m pytest -q --no-header -p no:cacheprovider"
run simt 600 $P tests/test_gpu_kernels.py -m gpu -k "csr or subgraph or spmm or row_kernels or dropout or pack"
run gemm_nt 600 $P tests/test_gpu_kernels.py -m gpu -k "gemm_nt"
run gemm_tn 600 $P tests/test_gpu_kernels.py -m gpu -k "gemm_tn"
run attn 600 $P tests/test_gpu_kernels.py -m gpu -k "views or partials"
run model 900 $P tests/test_gpu_model.py -m gpu
run smoke 300 python __graft_entry__.py smoke
run bench_tiny 300 python bench.py --workload tiny --steps 5 --warmup 3
run bench_products 900 python bench.py --steps 5 --warmup 3
cat $OUT/summary.txt
for f in simt gemm_nt gemm_tn attn model; do echo "== $f"; tail -n 12 $OUT/$f.log; done
echo "== smoke"; tail -n 5 $OUT/smoke.log
echo "== bench_tiny"; tail -n 3 $OUT/bench_tiny.log
echo "== bench_products"; tail -n 3 $OUT/bench_products.log
