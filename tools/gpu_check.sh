#!/bin/bash
# One single-GPU visit: smoke, GPU parity tests, kernel sweeps, headline bench (both arms), ncu launch list of the SAME
# bench command + one full capture of the two matrix kernels. Every step logs into gpurun_out/ and records its exit
# status; later steps run even if an earlier one fails. Knobs: SKIP_SWEEP=1, SKIP_NCU=1, BENCH_STEPS, BENCH_WARMUP.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
STATUS=$OUT/status.txt
: > $STATUS
step() { local name=$1; shift; ( "$@" ) > $OUT/$name.log 2>&1; local rc=$?; echo "$name rc=$rc" | tee -a $STATUS; return $rc; }

nvidia-smi > $OUT/nvidia-smi.txt 2>&1
( nproc; lscpu | grep -E "Model name|Socket|Core|Thread|^CPU\(s\)"; free -g | head -2 ) > $OUT/host.txt 2>&1

step smoke timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
step pytest_gpu timeout 1500 python -m pytest tests -m gpu -q --timeout 600
tail -3 $OUT/pytest_gpu.log
if [ "${SKIP_SWEEP:-0}" != "1" ]; then
  step sweep timeout 900 python tools/sweep.py --reps 10 --sustained 200
  step sweep_f32 timeout 900 python tools/sweep.py --storage f32 --reps 10
  grep -h BEST $OUT/sweep.log $OUT/sweep_f32.log | cut -c1-300
fi
BENCH="python bench.py --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-3}"
step bench timeout 1200 $BENCH
step bench_ref timeout 900 python bench.py --impl reference --steps 2 --warmup 1
step bench_f32 timeout 900 python bench.py --storage f32 --no-cpu-baseline
if [ "${SKIP_NCU:-0}" != "1" ]; then
  # launch list of the same command as the bench line above (cold-cache, serialised: compare SHARES)
  step ncu_plain timeout 900 $BENCH --no-cpu-baseline && \
  step ncu_launches timeout 2400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_bench_default.csv $BENCH --no-cpu-baseline
  SMALL="python bench.py --N 20000 --Mt 106250 --steps 1 --warmup 1 --no-cpu-baseline"
  step ncu_small_plain timeout 600 $SMALL && \
  step ncu_full timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_ax_partial|k_atx_cta' -s 40 -c 4 -f -o $OUT/prof_matrix $SMALL
fi
cat $STATUS
