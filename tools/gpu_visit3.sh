#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
timeout 600 python tools/sweep.py --storage f32 --reps 20 > $OUT/sweep_f32.log 2>&1; echo "sweep f32 rc=$?"; grep -E "BEST" $OUT/sweep_f32.log | cut -c1-400
timeout 600 python tools/sweep.py --quick --sustained 200 > $OUT/sweep_f64_check.log 2>&1; echo "sweep f64 rc=$?"; grep -E "default" $OUT/sweep_f64_check.log | cut -c1-200
timeout 900 python bench.py --storage f32 --no-cpu-baseline > $OUT/bench_f32.log 2>&1; echo "bench f32 rc=$?"; tail -1 $OUT/bench_f32.log | cut -c1-300
