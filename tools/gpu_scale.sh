#!/bin/bash
# Multi-GPU visit (gpurun --gpus G): shard-invariance tests through main_meth --gpus, then bench.py at the rank counts
# in SCALE_LIST, each with the fused peer-memory all-reduce (default) and, if AB=1, again with NCCL collectives.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
G=${1:-2}
nvidia-smi topo -m > $OUT/topo_$G.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > $OUT/pytest_multi_$G.log 2>&1; echo "pytest_multi rc=$?" | tee $OUT/status_scale_$G.txt
tail -5 $OUT/pytest_multi_$G.log
run_bench() {  # n tag env
  local n=$1 tag=$2
  if [ "$n" = "1" ]; then
    timeout 1200 python bench.py --gpus 1 --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-3} --no-cpu-baseline > $OUT/bench_g${n}_of${G}${tag}.log 2>&1
  else
    NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-3} > $OUT/bench_g${n}_of${G}${tag}.log 2>&1
  fi
  echo "bench n=$n$tag rc=$?" | tee -a $OUT/status_scale_$G.txt
  grep -h '^{' $OUT/bench_g${n}_of${G}${tag}.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print(d['n_gpus'], 'it/s', round(d['value'], 3), 'ms', round(d['ms_per_step'], 2), d['config'].get('cross_gpu_sums'), {k: round(v, 3) for k, v in r['phase_ms_per_step'].items()})
" || tail -20 $OUT/bench_g${n}_of${G}${tag}.log
}
for n in ${SCALE_LIST:-1 $G}; do
  run_bench $n ""
  if [ "${AB:-0}" = "1" ] && [ "$n" != "1" ]; then VAMPOMI_XCHG=0 run_bench $n "_nccl"; fi
done
if [ "${F32:-0}" = "1" ]; then   # opt-in FP32-storage mode at full width, for the record (not the headline)
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $G --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-3} --storage f32 > $OUT/bench_g${G}_of${G}_f32.log 2>&1
  echo "bench n=$G f32 rc=$?" | tee -a $OUT/status_scale_$G.txt
  grep -h '^{' $OUT/bench_g${G}_of${G}_f32.log | tail -1 | cut -c1-200
fi
