"""Wall-clock breakdown of one reconcile step through the binding (host overheads vs device phases)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding, abi
import torch

img, c = wva.synth.baseline_config(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
ctx = binding.Context(0)
if len(sys.argv) > 2:
    ctx.grid_set_fused(int(sys.argv[2]))
ctx.upload(img)
R, B = c["r_max"], c["b_max"]
def step():
    t0 = time.perf_counter(); ctx.analyze(R, B, want_cube=True)
    t1 = time.perf_counter(); ctx.solve(unlimited=True, download=False)
    t2 = time.perf_counter(); ctx.allocate_by_type()
    t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3
for _ in range(5): step()
rows = np.array([step() for _ in range(30)])
print("wall ms  analyze %.3f  solve %.3f  totals %.3f  | sum %.3f" % (*np.median(rows, axis=0), np.median(rows.sum(axis=1))))
print("device phases us: pairs %d grid %d (kernel %d heavy %d) solve %d totals %d" % tuple(ctx.phase_usec(p) for p in (abi.PHASE_PAIRS, abi.PHASE_GRID, abi.PHASE_GRID_KERNEL, abi.PHASE_GRID_HEAVY, abi.PHASE_SOLVE, abi.PHASE_TOTALS)))
# pairs alone, grid alone
ts = []
for _ in range(10):
    t0 = time.perf_counter(); ctx.analyze_pairs(download=False); ts.append((time.perf_counter() - t0) * 1e3)
print("analyze_pairs alone wall ms %.3f (device %d us)" % (np.median(ts), ctx.phase_usec(abi.PHASE_PAIRS)))
ts = []
for _ in range(10):
    t0 = time.perf_counter(); ctx.analyze_grid_device(R, B, want_cube=True); ts.append((time.perf_counter() - t0) * 1e3)
print("analyze_grid alone wall ms %.3f (device %d us, kernel %d)" % (np.median(ts), ctx.phase_usec(abi.PHASE_GRID), ctx.phase_usec(abi.PHASE_GRID_KERNEL)))
