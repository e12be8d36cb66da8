#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 900 python tools/sweep.py --quick --sustained ${SUSTAINED:-300} > $OUT/sweep_sustained.log 2>&1; echo "sweep rc=$?"
tail -30 $OUT/sweep_sustained.log
