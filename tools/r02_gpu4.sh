#!/bin/bash
# round 2, GPU pass 4: tuned sweep kernels (hoisted epilogue, lean on its own stream), greedy prefetch rings, host mirror deltas/adapters
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/grid_ab.py 3 1 33 17 > gpurun_out/r02h_grid_ab_cfg3.txt 2>&1
timeout 300 python tools/grid_ab.py 2 1 17 > gpurun_out/r02h_grid_ab_cfg2.txt 2>&1
timeout 600 python tools/greedy_stats.py > gpurun_out/r02h_greedy_stats.txt 2>&1
( time timeout 1800 python -m pytest tests -m gpu -x -q -s --durations=8 ) > gpurun_out/r02h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err
timeout 600 python bench.py --config 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_n1_cfg2.json 2> gpurun_out/r02h_bench_n1_cfg2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02h_launches_cfg3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan_ -s 9 -c 3 -o gpurun_out/r02h_scan \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_ncu_scan.log 2>&1
ls -la gpurun_out | tail -12
