"""Latency of one CreateAllocation (BASELINE config 1) through the warp-per-pair kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding
ctx = binding.Context(0)
for rpm, mb in [(600.0, 512), (6000.0, 512), (600.0, 64), (600.0, 8)]:
    img = wva.synth.config1(); img.srv_arrival_rpm[0] = rpm; img.srv_max_batch[0] = mb
    ctx.upload(img)
    ts = []
    for i in range(5):
        ctx.analyze_pairs(download=False); ts.append(ctx.phase_usec(wva.abi.PHASE_PAIRS))
    print("rpm %6.0f N %3d: usec min %d  %s" % (rpm, mb, min(ts), ctx.pair_counters()), flush=True)
