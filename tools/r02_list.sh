#!/bin/bash
# exact-chain list phase: A/B of the sweep + the parity tests that exercise the deferred chains (1 GPU)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-r02l}
timeout 300 python tools/grid_ab.py 3 1 > gpurun_out/${T}_grid_ab_cfg3.txt 2>&1
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_full_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
cat gpurun_out/${T}_grid_ab_cfg3.txt; tail -4 gpurun_out/${T}_pytest.log; cut -c1-300 gpurun_out/${T}_bench_cfg3.json; python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_cfg3.json").read().strip().splitlines()[-1])
print(d.get("phases_ms"), d.get("e2e"))
PY
