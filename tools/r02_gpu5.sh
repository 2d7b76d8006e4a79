#!/bin/bash
# round 2, pass 5 (2 GPUs): reordered sweep streams, greedy, all GPU tests incl. the 2-rank / 2-device ones, N = 1, 2 bench lines
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/grid_ab.py 3 1 17 > gpurun_out/r02e_grid_ab_cfg3.txt 2>&1
timeout 300 python tools/grid_ab.py 2 1 17 > gpurun_out/r02e_grid_ab_cfg2.txt 2>&1
timeout 600 python tools/greedy_stats.py > gpurun_out/r02e_greedy_stats.txt 2>&1
( time timeout 1800 python -m pytest tests -m gpu -x -q -s --durations=8 ) > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
bash tools/r02_multi.sh "1 2" r02e
