# r2l (1 GPU): ncu --set full of the BatchNorm backward launches and one SpMM launch of the final tree
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --no-cpu-baseline --no-e2e --no-extra --no-graph --steps 1 --warmup 3"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_kernel -s 24 -c 8 -o $OUT/r2l_bn_bwd $B > /dev/null 2>&1; echo "ncu bn_bwd rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_kernel -s 18 -c 2 -o $OUT/r2l_spmm $B > /dev/null 2>&1; echo "ncu spmm rc=$?"
ls -la $OUT/r2l_*
