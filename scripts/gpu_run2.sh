set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
P="python -m pytest -q --no-header -p no:cacheprovider"
timeout 600 $P tests/test_gpu_kernels.py -m gpu -k "csr or subgraph or spmm or row_kernels or dropout or pack" > $OUT/simt.log 2>&1; echo "simt rc=$?"
timeout 900 $P tests/test_gpu_model.py -m gpu > $OUT/model.log 2>&1; echo "model rc=$?"
timeout 300 python bench.py --workload arxiv --steps 5 --warmup 3 > $OUT/bench_arxiv.log 2>&1; echo "arxiv rc=$?"
timeout 300 python bench.py --workload pokec --steps 5 --warmup 3 > $OUT/bench_pokec.log 2>&1; echo "pokec rc=$?"
bash scripts/gpu_profile.sh products
tail -n 15 $OUT/simt.log; tail -n 40 $OUT/model.log
