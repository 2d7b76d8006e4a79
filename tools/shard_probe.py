"""Why is one shard of the 8-GPU weak-scaling job slower?  Sweep alone, shard `g`, fused and stop-and-go flows (GPU).
usage: python tools/shard_probe.py [shard=5] [world=8]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wva_import
wva = wva_import.load()
from inferno_autoscaler_b200 import binding, abi

g = int(sys.argv[1]) if len(sys.argv) > 1 else 5
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
c = dict(wva.synth.CONFIGS[3])
img = wva.synth.make_system(c["S"] * world, c["A"], seed=3, n_types=c["T"])
ctx = binding.Context(0)
ctx.upload(img)
R, B = c["r_max"], c["b_max"]
for shard in (g, 0):
    first = img.S * shard // world
    ctx.set_shard(first, img.S // world)
    for fused in (1, 0):
        ctx.grid_set_fused(fused)
        t = []
        for i in range(6):
            ctx.analyze_grid_device(R, B, want_cube=True)
            t.append((ctx.phase_usec(abi.PHASE_GRID), ctx.phase_usec(abi.PHASE_GRID_KERNEL), ctx.phase_usec(abi.PHASE_GRID_HEAVY)))
        t = np.array(t[2:], dtype=float) / 1e3
        print(json.dumps({"shard": shard, "fused_asked": fused, "fused": ctx.grid_last_fused(), "grid_ms": round(float(t[:, 0].mean()), 3),
                          "sweep_kernel_ms": round(float(t[:, 1].mean()), 3), "exact_chain_ms": round(float(t[:, 2].mean()), 3),
                          "lists": ctx.grid_list_sizes()}))
    # the deferred candidates: batch sizes and replicas
    ids, n = ctx.grid_deferred(cap=1 << 22)
    ids = ids.astype(np.int64)
    b = ids % B + 1
    r = (ids // B) % R + 1
    pair = ids // (B * R)
    srv = pair // img.A + first
    u, cnt = np.unique(srv, return_counts=True)
    top = np.argsort(-cnt)[:5]
    print(json.dumps({"shard": shard, "deferred": int(n), "b_mean": float(b.mean()), "b_max": int(b.max()), "share_b_ge_480": float((b >= 480).mean()),
                      "servers_with_deferred": int(len(u)), "top_servers": [(int(u[i]), int(cnt[i])) for i in top],
                      "top_server_params": [dict(arr=float(img.srv_arrival_rpm[u[i]]), out=int(img.srv_out_tokens[u[i]]), inp=int(img.srv_in_tokens[u[i]])) for i in top]}))
