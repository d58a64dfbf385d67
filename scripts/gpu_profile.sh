#!/bin/bash
# ncu evidence on the GPU box: launch list of the bench step + full captures of the dominant kernels.
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
WL=${1:-products}
B="python bench.py --workload $WL --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_$WL.csv $B --steps 2 --warmup 3 > $OUT/ncu_launches_$WL.log 2>&1
echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 7 -c 2 -f -o $OUT/spmm_full_$WL $B --steps 1 --warmup 3 > $OUT/ncu_spmm_$WL.log 2>&1
echo "spmm full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 20 -c 3 -f -o $OUT/gemm_nt_full_$WL $B --steps 1 --warmup 3 > $OUT/ncu_gemm_nt_$WL.log 2>&1
echo "gemm_nt full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 10 -c 2 -f -o $OUT/gemm_tn_full_$WL $B --steps 1 --warmup 3 > $OUT/ncu_gemm_tn_$WL.log 2>&1
echo "gemm_tn full rc=$?"
ls -la $OUT | tail -20
