#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
timeout 600 python tools/sweep.py --storage f32 --quick --sustained 200 > $OUT/sweep_f32.log 2>&1; echo "sweep f32 rc=$?"; grep -E "default|BEST" $OUT/sweep_f32.log | cut -c1-200
timeout 900 python bench.py --storage f32 --no-cpu-baseline > $OUT/bench_f32.log 2>&1; echo "bench f32 rc=$?"; tail -1 $OUT/bench_f32.log | cut -c1-400
SMALL="python bench.py --N 20000 --Mt 106250 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $SMALL > $OUT/ncu_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv $SMALL > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_ax_partial|k_atx_cta' -s 40 -c 4 -f -o $OUT/prof_matrix $SMALL > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
