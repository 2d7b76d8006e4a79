"""A/B of the sweep kernels on one workload (GPU): device time of wva_analyze_grid_device per mode.
usage: python tools/grid_ab.py [config] [modes...]   (modes are wva_set_certified_tails values; default 1 33 17)"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
modes = [int(x) for x in sys.argv[2:]] or [1, 33, 17]      # 101 = mode 1 with the fused flow switched off (stop and go)
img, c = wva.synth.baseline_config(cfg)
ctx = binding.Context(0)
ref = None
for mode in modes:
    ctx.set_certified_tails(mode % 100)
    ctx.grid_set_fused(mode < 100)
    ctx.upload(img)
    t = []
    for i in range(6):
        ctx.analyze_grid_device(c["r_max"], c["b_max"], want_cube=True)
        t.append((ctx.phase_usec(wva.abi.PHASE_GRID), ctx.phase_usec(wva.abi.PHASE_GRID_KERNEL), ctx.phase_usec(wva.abi.PHASE_GRID_HEAVY)))
    best = ctx.grid_fetch()
    if ref is None:
        ref = best.tobytes()
    t = np.array(t[2:], dtype=float) / 1e3
    print(json.dumps({"config": cfg, "mode": mode, "grid_ms": round(float(t[:, 0].mean()), 3), "sweep_kernel_ms": round(float(t[:, 1].mean()), 3),
                      "exact_chain_ms": round(float(t[:, 2].mean()), 3), "fused": ctx.grid_last_fused(), "lists": ctx.grid_list_sizes(), "same_winners": best.tobytes() == ref,
                      "counters": ctx.grid_counters()}))
ctx.close()
