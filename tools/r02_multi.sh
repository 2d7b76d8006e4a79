#!/bin/bash
# round 2 multi-GPU pass (gpurun --gpus N): library collectives (wva_comm_*, wva_group_*), weak scaling of the default
# workload, strong scaling of BASELINE configs 4 / 5 with the in-run check that sharded decisions equal a 1-rank pass.
# usage: bash tools/r02_multi.sh "<list of N>" <tag>
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NS=${1:-"2"}; TAG=${2:-r02m}
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s ) > gpurun_out/${TAG}_pytest_multi.log 2>&1
run() {  # N, name, extra args
  local n=$1 name=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --no-cpu-baseline "$@" > gpurun_out/${TAG}_${name}_n1.json 2> gpurun_out/${TAG}_${name}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
        bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/${TAG}_${name}_n${n}.json 2> gpurun_out/${TAG}_${name}_n${n}.err
  fi
}
for n in $NS; do
  run $n weak3 --steps 20 --warmup 5
  run $n strong4 --config 4 --strong --verify --steps 10 --warmup 3
  run $n strong4lim --config 4 --strong --limited --verify --steps 10 --warmup 3
done
# config 5 (100 000 servers x 16 accelerators): the largest N only, plus N = 1 when asked for
LAST=$(echo $NS | awk '{print $NF}')
run $LAST strong5 --config 5 --strong --no-cube --steps 5 --warmup 3
ls -la gpurun_out | grep ${TAG}
