# r2f (1 GPU): bucketed CSR build: parity with the direct build, timing, per-kernel breakdown; e2e with it
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -q -m gpu -k "csr or symmetry or feeder or subset or subgraph or fullsize or model" -x > $OUT/r2f_tests.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" $OUT/r2f_tests.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/r2f_tests.log | cut -c1-300 | head -20
for B in 0 1; do SGF_CSR_BUCKETS=$B timeout 200 python scripts/bench_csr.py 2>&1 | tail -n 1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bucket|csr_|scan_" -c 40 --csv --log-file $OUT/r2f_csr_times.csv python scripts/bench_csr.py > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2f_csr_times.csv',errors='ignore')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
for r in rows[hi+1:hi+12]:
    if len(r)>mv: print(r[kn][:40], round(float(r[mv].replace(',',''))/1e6,3))
PY
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $OUT/bench_r2f.log 2>&1; echo "bench rc=$?"
grep "^{" $OUT/bench_r2f.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms/step', d['ms_per_step'], 'e2e', d['e2e'])"
