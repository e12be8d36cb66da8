#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/sanitize_case.py > $OUT/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 $OUT/sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_case.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/sanitize_memcheck.log
timeout 900 python bench.py --storage f32 --no-cpu-baseline > $OUT/bench_f32.log 2>&1; echo "bench f32 rc=$?"; tail -1 $OUT/bench_f32.log | cut -c1-300
