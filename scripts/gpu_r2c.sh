# r2c (1 GPU): GPU tests, the default bench line, and ncu --set full of the Gram-form attention kernels + launch list (profiles/r2_*)
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > $OUT/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed" $OUT/all_gpu.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/all_gpu.log | cut -c1-300 | head -30
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_default.log 2>&1; echo "bench rc=$?"
grep "^{" $OUT/bench_default.log | cut -c1-4000
# the vectorised operand-pack kernel (fp32 / bf16x3 path), never run on a GPU in r1: validate it and measure it
SGF_PACK_VEC=1 timeout 900 python -m pytest tests -q -m gpu -k "pack_operand or gemm_nt or gemm_tn or golden or arxiv or midsize or gram" > $OUT/pack_vec_tests.log 2>&1; echo "pytest with SGF_PACK_VEC=1 rc=$?"
grep -E "passed|failed" $OUT/pack_vec_tests.log
for PV in 0 1; do
  SGF_PACK_VEC=$PV timeout 300 python bench.py --workload arxiv --no-cpu-baseline --no-e2e --no-extra --steps 30 --warmup 5 > $OUT/bench_arxiv_pv$PV.log 2>&1
  grep "^{" $OUT/bench_arxiv_pv$PV.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('arxiv fp32 SGF_PACK_VEC=$PV ms/step', d['ms_per_step'])"
done
B="python bench.py --no-cpu-baseline --no-e2e --no-extra --no-graph"
# skip the warm-up steps' launches: full metric set of every instance of the attention kernels in the timed step
for KN in gram_kernel ln_bwd_attn_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KN -s 3 -c 1 -f -o $OUT/r2_$KN $B --steps 1 --warmup 3 > $OUT/ncu_$KN.log 2>&1; echo "ncu $KN rc=$?"
done
# gemm_nt instances of one step in launch order (the ATTN_GRAM apply <65> and the dx GEMM <49> are among them)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 45 -c 15 -f -o $OUT/r2_gemm_nt $B --steps 1 --warmup 3 > $OUT/ncu_gemm_nt.log 2>&1; echo "ncu gemm_nt rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/r2_launches_products.csv $B --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "launches rc=$?"
ls -la $OUT/*.ncu-rep | tail
