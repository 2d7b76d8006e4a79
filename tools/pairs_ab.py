"""A/B timing of the pair-sizing kernels inside one process (same GPU, same clocks)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
img, c = wva.synth.baseline_config(cfg)
ctx = binding.Context(0)
ctx.upload(img)
for name, warp_max, knob, cert in [("warp certified", 1 << 22, 0, 1), ("warp exact", 1 << 22, 0, 0)]:
    ctx.pairs_set_warp_max(warp_max); ctx.pairs_set_pstore(knob); ctx.set_certified_tails(cert); ctx.upload(img)
    ts = []
    for i in range(6):
        ctx.analyze_pairs(download=False)
        ts.append(ctx.phase_usec(wva.abi.PHASE_PAIRS))
    print("%-18s pairs phase usec: min %d median %d  %s" % (name, min(ts), int(np.median(ts)), ctx.pair_counters()), flush=True)
