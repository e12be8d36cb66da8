#!/bin/bash
# General single-GPU visit: full GPU test suite, config timings, ingest bench, headline bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 900 python tools/run_configs.py > $OUT/configs.json 2> $OUT/configs.err; echo "configs rc=$?"; tail -3 $OUT/configs.err; grep -A12 C5_ $OUT/configs.json
timeout 900 python tools/load_bench.py > $OUT/load_bench.json 2> $OUT/load_bench.err; echo "load rc=$?"; cat $OUT/load_bench.json; tail -3 $OUT/load_bench.err
timeout 1200 python bench.py > $OUT/bench.log 2>&1; echo "bench rc=$?"; tail -1 $OUT/bench.log | cut -c1-300
