#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-r02u}
timeout 300 python tools/grid_ab.py 3 1 101 > gpurun_out/${T}_grid_ab_cfg3.txt 2>&1
timeout 300 python tools/grid_ab.py 4 1 101 > gpurun_out/${T}_grid_ab_cfg4.txt 2>&1
( timeout 900 python -m pytest tests/test_parity_full_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "fused or overlapped or config4 or unlimited" ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
cat gpurun_out/${T}_grid_ab_cfg3.txt gpurun_out/${T}_grid_ab_cfg4.txt | cut -c1-230; tail -4 gpurun_out/${T}_pytest.log; python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_cfg3.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d.get("phases_ms"), d.get("e2e",{}).get("ms_per_step"))
PY
