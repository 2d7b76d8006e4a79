#!/bin/bash
# round 2, 2-GPU pass: weak scaling point N=2 of the default workload, config 5 on one GPU (denominator of its strong scaling)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r02m2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 \
    bench.py --gpus 2 --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/${TAG}_weak3_n2.json 2> gpurun_out/${TAG}_weak3_n2.err
timeout 300 python bench.py --gpus 1 --no-cpu-baseline --config 5 --strong --no-cube --steps 3 --warmup 3 > gpurun_out/${TAG}_strong5_n1.json 2> gpurun_out/${TAG}_strong5_n1.err
ls -la gpurun_out | grep ${TAG}
