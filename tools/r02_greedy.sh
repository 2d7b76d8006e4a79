#!/bin/bash
# greedy solver paths: parity tests + instrumentation (1 GPU)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-r02g}
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "greedy or solve" ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python tools/greedy_stats.py 10000 2 > gpurun_out/${T}_greedy_stats.txt 2>&1
timeout 300 python bench.py --config 4 --limited --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_cfg4_limited.json 2> gpurun_out/${T}_cfg4_limited.err
tail -5 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_greedy_stats.txt; cat gpurun_out/${T}_cfg4_limited.json | cut -c1-400
