# quick GPU check of a subset of tests: bash scripts/gpu_quick.sh "<pytest -k expression>"
set +e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -k "$1" > $OUT/quick.log 2>&1; echo "rc=$?"
grep -E "passed|failed" $OUT/quick.log; grep -E "^(FAILED|E   [A-Za-z])" $OUT/quick.log | cut -c1-400 | head -20
