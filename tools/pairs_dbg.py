import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
A = int(sys.argv[2]) if len(sys.argv) > 2 else 4
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 2
img = wva.synth.make_system(S, A, seed=seed, n_types=A)
ctx = binding.Context(0)
ctx.upload(img); ctx.pairs_set_pstore(4)
for i in range(2): ctx.analyze_pairs(download=False)
d = ctx.pair_debug()
w0 = d[:, 0]; cyc = (w0 & np.uint64(0xfffffffff)).astype(np.int64); pend = ((w0 >> np.uint64(36)) & np.uint64(0xff)).astype(np.int64); ksteps = (w0 >> np.uint64(44)).astype(np.int64); tot = (d[:, 1] >> np.uint64(32)).astype(np.int64); act = (d[:, 1] & np.uint64(0xffffffff)).astype(np.int64)
order = np.argsort(-cyc)[:12]
for p in order:
    s, a = p // img.A, p % img.A
    N = img.srv_max_batch[s] if img.srv_max_batch[s] > 0 else max(1, img.perf_max_batch[img.srv_model[s]*img.A+a] * img.perf_at_tokens[img.srv_model[s]*img.A+a] // max(1, img.srv_out_tokens[s]))
    print("pair %3d cycles %9d (%.2f ms) rounds %2d active %2d pend %d ksteps %d  N %4d slo_ttft %g itl %g tps %g rpm %g" % (p, cyc[p], cyc[p]/1.965e6, tot[p], act[p], pend[p], ksteps[p], N, img.srv_slo_ttft[s], img.srv_slo_itl[s], img.srv_slo_tps[s], img.srv_arrival_rpm[s]))
print("pairs", len(cyc), "cycles percentiles 50/90/99/max", [int(np.percentile(cyc, q)) for q in (50, 90, 99, 100)], "sum Mcyc %.1f" % (cyc.sum() / 1e6))
print("phase usec", ctx.phase_usec(wva.abi.PHASE_PAIRS), "sum active rounds", act.sum(), "max", act.max())
