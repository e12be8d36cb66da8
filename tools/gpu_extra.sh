#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python tools/run_configs.py > $OUT/configs.json 2> $OUT/configs.err; echo "configs rc=$?"; tail -5 $OUT/configs.err
timeout 1500 python tests/tools/parity_c2.py > $OUT/parity_c2.json 2> $OUT/parity_c2.err; echo "parity rc=$?"; tail -5 $OUT/parity_c2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/parity_c2.json"))
    for r in d["runs"]:
        print(r["case"], "pass", r["pass"], "max dev", max(max(x) for x in r["rel_l2_x1_r1_per_iteration"]), "gpu it s", r["gpu_iter_s"], "ref it s", r["ref_iter_s"], "load", r["gpu_load_s"])
except Exception as e:
    print("no parity json", e)
PY
