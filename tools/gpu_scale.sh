#!/bin/bash
# Multi-GPU visit (gpurun --gpus G): shard-invariance tests through main_meth --gpus, then bench.py at 1..G ranks.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out
mkdir -p $OUT
G=${1:-2}
nvidia-smi topo -m > $OUT/topo_$G.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > $OUT/pytest_multi_$G.log 2>&1; echo "pytest_multi rc=$?" | tee $OUT/status_scale_$G.txt
for n in ${SCALE_LIST:-1 $G}; do
  if [ "$n" = "1" ]; then
    timeout 1200 python bench.py --gpus 1 --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-3} --no-cpu-baseline > $OUT/bench_g1_of$G.log 2>&1
  else
    NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-3} > $OUT/bench_g${n}_of$G.log 2>&1
  fi
  echo "bench n=$n rc=$?" | tee -a $OUT/status_scale_$G.txt
  grep -h '^{' $OUT/bench_g${n}_of$G.log | tail -1 | cut -c1-400
done
