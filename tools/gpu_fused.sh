#!/bin/bash
# Single-GPU visit for the fused schedule: parity tests, sweep of the multi-vector kernels, bench A/B (fused vs plain).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
STATUS=$OUT/status_fused.txt
: > $STATUS
step() { local name=$1; shift; ( "$@" ) > $OUT/$name.log 2>&1; local rc=$?; echo "$name rc=$rc" | tee -a $STATUS; return $rc; }
step smoke timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
step pytest_gpu timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x
tail -15 $OUT/pytest_gpu.log
step sweep_multi timeout 600 python tools/sweep.py --multi ${MULTI_REPS:-40}
grep -h BEST $OUT/sweep_multi.log | cut -c1-1500
step bench_fused timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
tail -c 3000 $OUT/bench_fused.log
step bench_plain timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --schedule plain
tail -c 1500 $OUT/bench_plain.log
cat $STATUS
