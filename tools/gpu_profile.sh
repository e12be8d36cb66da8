#!/bin/bash
# One single-GPU visit that produces everything profiles/ cites for the current default (onepass schedule): smoke, GPU parity
# tests, parity reports against the reference fixtures per schedule, headline bench (both arms), ncu launch list of the
# bench command and one full capture of the two dominant kernels on one 8-GPU shard of the headline configuration.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
STATUS=$OUT/status_profile.txt
: > $STATUS
step() { local name=$1; shift; ( "$@" ) > $OUT/$name.log 2>&1; local rc=$?; echo "$name rc=$rc" | tee -a $STATUS; return $rc; }
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
( nproc; lscpu | grep -E "Model name|Socket|Core|Thread|^CPU\(s\)"; free -g | head -2 ) > $OUT/host.txt 2>&1
step smoke timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
step pytest_gpu timeout 1500 python -m pytest tests -m gpu -q --timeout 600
tail -3 $OUT/pytest_gpu.log
for s in onepass recycled fused plain; do
  timeout 600 python tests/tools/parity_report.py $s > $OUT/parity_report_$s.json 2> $OUT/parity_report_$s.err; echo "parity_report_$s rc=$?" | tee -a $STATUS
done
BENCH="python bench.py --steps ${BENCH_STEPS:-5} --warmup ${BENCH_WARMUP:-3}"
step bench timeout 1200 $BENCH
step bench_ref timeout 900 python bench.py --impl reference --steps ${BENCH_STEPS:-5} --warmup ${BENCH_WARMUP:-3}
if [ "${SKIP_C2:-0}" != "1" ]; then
  timeout 1500 python tests/tools/parity_c2.py > $OUT/parity_c2.json 2> $OUT/parity_c2.err; echo "parity_c2 rc=$?" | tee -a $STATUS
fi
if [ "${SKIP_CONFIGS:-0}" != "1" ]; then
  timeout 900 python tools/run_configs.py > $OUT/configs.json 2> $OUT/configs.err; echo "configs rc=$?" | tee -a $STATUS
fi
if [ "${ONLY_NCU:-0}" == "1" ]; then :; fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  NB="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ab --no-parity"
  step ncu_plain timeout 900 $NB && \
  step ncu_launches timeout 2400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_bench_default.csv $NB
  SMALL="python bench.py --N 20000 --Mt 106250 --steps 1 --warmup 1 --no-cpu-baseline --no-ab --no-parity"
  step ncu_small_plain timeout 600 $SMALL && \
  step ncu_full timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_gram|k_ax_multi' -s 6 -c 4 -f -o $OUT/prof_multi $SMALL
fi
cat $STATUS
