#!/bin/bash
# round 2, 8-GPU pass: weak scaling of the default workload (config 3 per GPU), strong scaling of BASELINE configs 4 (unlimited and
# capacity-limited) and 5 with the in-run check that sharded decisions equal a one-rank pass.
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r02f}
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
run() {  # N, name, extra args
  local n=$1 name=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline "$@" > gpurun_out/${TAG}_${name}_n1.json 2> gpurun_out/${TAG}_${name}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
        bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/${TAG}_${name}_n${n}.json 2> gpurun_out/${TAG}_${name}_n${n}.err
  fi
}
if [ "${2:-full}" = "full" ]; then
  ( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/${TAG}_pytest_multi.log 2>&1
  for n in 2 4; do run $n weak3 --steps 20 --warmup 5; done
  run 1 strong5 --config 5 --strong --no-cube --steps 3 --warmup 3
fi
run 8 weak3 --steps 20 --warmup 5
run 8 strong4 --config 4 --strong --verify --steps 10 --warmup 3
run 8 strong4lim --config 4 --strong --limited --verify --steps 10 --warmup 3
run 8 strong5 --config 5 --strong --no-cube --verify --steps 5 --warmup 3
ls -la gpurun_out | grep ${TAG}
