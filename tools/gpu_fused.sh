#!/bin/bash
# Single-GPU visit for the fused / recycled schedules: parity tests, sweep of the multi-vector kernels, bench A/B.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
STATUS=$OUT/status_fused.txt
: > $STATUS
step() { local name=$1; shift; ( "$@" ) > $OUT/$name.log 2>&1; local rc=$?; echo "$name rc=$rc" | tee -a $STATUS; return $rc; }
step smoke timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
step pytest_gpu timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ${PYTEST_ARGS:-}
tail -15 $OUT/pytest_gpu.log
if [ "${SKIP_SWEEP:-0}" != "1" ]; then
  step sweep_multi timeout 600 python tools/sweep.py --multi ${MULTI_REPS:-40}
  grep -h BEST $OUT/sweep_multi.log | cut -c1-1500
fi
for sched in ${SCHEDULES:-recycled fused plain}; do
  step bench_$sched timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 --no-cpu-baseline --schedule $sched
  grep -h '^{' $OUT/bench_$sched.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print(d['config']['schedule'], 'it/s', round(d['value'], 3), 'e2e', round(d['e2e']['value'], 3), 'ms', round(d['ms_per_step'], 2), r['kernel'], round(r['achieved']), {k: round(v, 3) for k, v in r['phase_ms_per_step'].items()}, r['whole_iteration']['passes'], d['config']['cg_iters_per_step'])
" || tail -20 $OUT/bench_$sched.log
done
cat $STATUS
