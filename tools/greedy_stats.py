"""Instrumentation run of the limited-capacity solve: pops, failed placements, cycles (profiling aid)."""
import sys, time
import numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wva_import
wva_import.load()
from inferno_autoscaler_b200 import synth, binding, abi

S = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
img, c = synth.baseline_config(4, n_servers=S)
ctx = binding.Context(0)
ctx.upload(img)
ctx.analyze_pairs(download=False)
acc, ch = ctx.solve(unlimited=True)
synth.set_capacity_from_demand(img, ch.acc, ch.num_replicas, fraction=0.6)
RANKED = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx.solve_set_ranked(RANKED)
for policy in (abi.POLICY_PRIORITY_EXHAUSTIVE, abi.POLICY_NONE, abi.POLICY_ROUND_ROBIN):
    for delayed in (False, True):
        ctx.upload(img)
        ctx.analyze_pairs(download=False)
        ctx.solve(unlimited=False, policy=policy, delayed_best_effort=delayed, download=False)
        st = ctx.solve_stats()
        path = ctx.solve_greedy_path()
        ms = ctx.phase_usec(abi.PHASE_SOLVE) / 1e3
        if path == 3:
            print("policy", policy, "delayed", delayed, "path", path, "solve ms %.3f" % ms, "events", st[0] & 0xffffffff, "groups with a stack", st[0] >> 32,
                  "batches", st[1] & 0xffffffff, "runs popped", st[1] >> 32, "scan Mcyc %.2f" % (st[2] / 1e6), "bestEffort Mcyc %.2f" % (st[3] / 1e6))
        else:
            print("policy", policy, "delayed", delayed, "path", path, "solve ms %.3f" % ms,
                  "pops", st[0], "fails", st[1], "queue cycles/pop %.0f" % (st[2] / max(st[0], 1)), "queue Mcyc %.1f" % (st[2] / 1e6),
                  "bestEffort Mcyc %.1f" % (st[3] / 1e6), "raw", st)
