"""Per-shard step time of the 8-GPU weak-scaling workload, measured on ONE GPU: the job's image (8 x config 3, as
`bench.py --gpus 8` generates it) is swept shard by shard -- what each rank of the 8-GPU run does on its own GPU.
usage: python tools/shard_times.py [world=8]"""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding, abi
import torch

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
c = dict(wva.synth.CONFIGS[3])
img = wva.synth.make_system(c["S"] * world, c["A"], seed=3, n_types=c["T"])
ctx = binding.Context(0)
ctx.upload(img)
R, B = c["r_max"], c["b_max"]
rows = []
for g in range(world):
    first = img.S * g // world
    count = img.S * (g + 1) // world - first
    ctx.set_shard(first, count)
    def step():
        ctx.analyze(R, B, want_cube=True)
        ctx.solve(unlimited=True, download=False)
        ctx.allocate_by_type(download=False)
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    row = {"shard": g, "servers": [int(first), int(first + count)], "ms_per_step": round(ms, 3),
           "pairs_ms": ctx.phase_usec(abi.PHASE_PAIRS) / 1e3, "grid_ms": ctx.phase_usec(abi.PHASE_GRID) / 1e3,
           "exact_chain_ms": ctx.phase_usec(abi.PHASE_GRID_HEAVY) / 1e3, "deferred": ctx.grid_list_sizes()["deferred"]}
    rows.append(row)
    print(json.dumps(row))
ms = [r["ms_per_step"] for r in rows]
print(json.dumps({"max_ms": max(ms), "mean_ms": round(float(np.mean(ms)), 3), "min_ms": min(ms)}))
