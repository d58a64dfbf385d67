# r2h (1 GPU): direct CSR build with batched loads: parity tests, window sweep, per-kernel times, e2e
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -q -m gpu -k "csr or symmetry or feeder or subset or subgraph or undirected or fullsize or model" -x > $OUT/r2h_tests.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" $OUT/r2h_tests.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/r2h_tests.log | cut -c1-300 | head -20
for W in 0 64 96 128 192; do SGF_CSR_FILL_WINDOW_MB=$W timeout 200 python scripts/bench_csr.py 2>&1 | tail -n 1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"csr_|scan_" -s 45 -c 15 --csv --log-file $OUT/r2h_csr_times.csv python scripts/bench_csr.py > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2h_csr_times.csv',errors='ignore')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
for r in rows[hi+1:hi+17]:
    if len(r)>mv: print(r[kn][:40], round(float(r[mv].replace(',',''))/1e6,3))
PY
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $OUT/bench_r2h.log 2>&1; echo "bench rc=$?"
grep "^{" $OUT/bench_r2h.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms/step', d['ms_per_step'], 'e2e', d['e2e'])"
