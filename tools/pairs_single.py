"""Latency of one CreateAllocation (BASELINE config 1) through the warp-per-pair kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding
ctx = binding.Context(0)
for cert in (1, 0):
    for knob in (4, 6):
        for rpm, mb in [(600.0, 512), (600.0, 64)]:
            img = wva.synth.config1(); img.srv_arrival_rpm[0] = rpm; img.srv_max_batch[0] = mb
            ctx.set_certified_tails(cert); ctx.pairs_set_pstore(knob)
            ctx.upload(img)
            ts = []
            for i in range(5):
                ctx.analyze_pairs(download=False); ts.append(ctx.phase_usec(wva.abi.PHASE_PAIRS))
            d = ctx.pair_debug()[0]
            print("cert %d %s N %3d: phase usec min %d kernel cycles %d (%.3f ms) rounds %d active %d %s" % (
                cert, "hbm " if knob & 2 else "smem", mb, min(ts), int(d[0]) & 0xfffffffff, (int(d[0]) & 0xfffffffff) / 1.965e6, int(d[1]) >> 32, int(d[1]) & 0xffffffff, ctx.pair_counters()), flush=True)
