#!/bin/bash
# round 2, evidence pass (1 GPU): every number quoted in README / DESIGN / profiles comes from this script's outputs.
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-r02z}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt; nproc >> gpurun_out/${T}_smi.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1_config3.json 2> gpurun_out/${T}_bench_n1_config3.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
timeout 600 python bench.py --config 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_n1_config2.json 2> /dev/null
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_n1_config4.json 2> /dev/null
timeout 600 python bench.py --config 4 --limited --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_n1_config4_limited.json 2> /dev/null
timeout 300 python tools/grid_ab.py 3 1 33 17 > gpurun_out/${T}_grid_ab_cfg3.txt 2>&1
timeout 300 python tools/greedy_stats.py 10000 2 > gpurun_out/${T}_greedy_stats.txt 2>&1
timeout 300 python tools/greedy_stats.py 10000 1 > gpurun_out/${T}_greedy_stats_ranked_queue.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_config3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scan_|k_grid_list|k_pairs_warp' -s 15 -c 5 -o gpurun_out/${T}_cfg3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_greedy_scan$|k_greedy_scan\(' -c 1 -o gpurun_out/${T}_greedy \
    python bench.py --config 4 --limited --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_greedy.log 2>&1
ls -la gpurun_out | grep ${T}
