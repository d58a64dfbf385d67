# the smaller BASELINE shapes on one GPU: bench numbers + launch lists (arxiv fp32, pokec bf16), preprocessing/eval timings, and
# ncu --set full captures of the attention GEMMs (QKV projection, attention apply) at the products shape
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
for WL in arxiv pokec; do
  timeout 300 python bench.py --workload $WL --no-cpu-baseline --no-e2e --steps 20 --warmup 5 > $OUT/bench_$WL.log 2>&1; echo "bench $WL rc=$?"
  grep "^{" $OUT/bench_$WL.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$WL ms_per_step', d['ms_per_step'], 'nodes/s', d['value'], 'spmm', d['roofline'] and d['roofline']['avg_launch_ms'])"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_$WL.csv python bench.py --workload $WL --no-cpu-baseline --no-e2e --no-graph --steps 2 --warmup 3 > $OUT/ncu_launches_$WL.log 2>&1; echo "launches $WL rc=$?"
done
timeout 600 python scripts/bench_prep.py --cpu 2>&1 | tee $OUT/prep_products.log
# gemm_nt captures: skip the launches of warm-up, take the forward QKV projection (768 wide) and the attention apply of a step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 57 -c 19 -f -o $OUT/gemm_nt_full_r1c python bench.py --no-cpu-baseline --no-e2e --no-graph --steps 1 --warmup 3 > $OUT/ncu_gemm_nt_r1c.log 2>&1; echo "gemm_nt full rc=$?"
ls -la $OUT/*.ncu-rep
