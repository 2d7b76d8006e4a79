"""Sweep kernel A/B inside one process: per-candidate / per-row (segmented) kernels, deferred-list sizes (profiling aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding, abi

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
img, c = wva.synth.baseline_config(cfg)
ctx = binding.Context(0)
ctx.upload(img)
for mode, name in ((3, "per candidate"), (5, "thread per row"), (9, "warp per row"), (1, "auto")):
    ctx.set_certified_tails(mode)
    ts = []
    for i in range(4):
        ctx.analyze_grid_device(c["r_max"], c["b_max"])
        ts.append((ctx.phase_usec(abi.PHASE_GRID), ctx.phase_usec(abi.PHASE_GRID_KERNEL), ctx.phase_usec(abi.PHASE_GRID_HEAVY)))
    print(name, "grid/light/heavy usec", min(ts), "lists (heavy, literal)", ctx.grid_list_sizes(), ctx.grid_counters())
ctx.set_certified_tails(1)
