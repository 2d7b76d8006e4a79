import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding
img, c = wva.synth.baseline_config(2)
ctx = binding.Context(0)
ctx.upload(img); ctx.pairs_set_pstore(4)
for i in range(2): ctx.analyze_pairs(download=False)
d = ctx.pair_debug()
cyc = d[:, 0].astype(np.int64); tot = (d[:, 1] >> np.uint64(32)).astype(np.int64); act = (d[:, 1] & np.uint64(0xffffffff)).astype(np.int64)
order = np.argsort(-cyc)[:12]
for p in order:
    s, a = p // img.A, p % img.A
    N = img.srv_max_batch[s] if img.srv_max_batch[s] > 0 else max(1, img.perf_max_batch[img.srv_model[s]*img.A+a] * img.perf_at_tokens[img.srv_model[s]*img.A+a] // max(1, img.srv_out_tokens[s]))
    print("pair %3d cycles %9d (%.2f ms) rounds %2d active %2d  N %4d slo_ttft %g itl %g tps %g rpm %g" % (p, cyc[p], cyc[p]/1.965e6, tot[p], act[p], N, img.srv_slo_ttft[s], img.srv_slo_itl[s], img.srv_slo_tps[s], img.srv_arrival_rpm[s]))
print("phase usec", ctx.phase_usec(wva.abi.PHASE_PAIRS), "sum active rounds", act.sum(), "max", act.max())
