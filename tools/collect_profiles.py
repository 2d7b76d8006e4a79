"""Copy the evidence of a GPU pass from gpurun_out/ (scratch) into profiles/ (tracked): bench JSON lines, the ncu launch list
and its per-kernel shares, key metrics of the `ncu --set full` captures.
usage: python tools/collect_profiles.py <tag in gpurun_out, e.g. r02z> [<multi-GPU tag, e.g. r02f>]"""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
mtag = sys.argv[2] if len(sys.argv) > 2 else None

def line(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None

for name in ("bench_n1_config3", "bench_reference_arm", "bench_n1_config2", "bench_n1_config4", "bench_n1_config4_limited", "bench_n1_config3_untamed"):
    d = line(os.path.join(G, "%s_%s.json" % (tag, name)))
    if d:
        json.dump(d, open(os.path.join(P, name.replace("bench_", "bench_r02_") + ".json"), "w"), indent=1)
for f in ("grid_ab_cfg3.txt", "greedy_stats.txt", "smi.txt"):
    src = os.path.join(G, "%s_%s" % (tag, f))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "r02_" + f))
src = os.path.join(G, "%s_launches_config3.csv" % tag)
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, "launches_r02_config3.csv"))
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) > mv:
            agg[r[kn].split("(")[0].replace("void ", "").replace("wva::", "")].append(float(r[mv]) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    shares = [{"kernel": k, "launches": len(v), "mean_us": round(sum(v) / len(v), 1), "share_pct": round(100 * sum(v) / tot, 2)}
              for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]
    json.dump({"source": "launches_r02_config3.csv (ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3; serialised, cold cache: compare shares)",
               "kernels": shares}, open(os.path.join(P, "launch_shares_r02.json"), "w"), indent=1)

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_active.max",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]

def extract(rep, out):
    if not os.path.exists(rep):
        return
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")], "report": os.path.basename(rep)}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (vals[i] + " " + units[i]).strip()
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(len(res), "kernels ->", out)

extract(os.path.join(G, "%s_cfg3.ncu-rep" % tag), os.path.join(P, "ncu_full_r02_cfg3.json"))
extract(os.path.join(G, "%s_greedy.ncu-rep" % tag), os.path.join(P, "ncu_full_r02_greedy.json"))
if mtag:
    for f in sorted(os.listdir(G)):
        if f.startswith(mtag + "_") and f.endswith(".json"):
            d = line(os.path.join(G, f))
            if d:
                json.dump(d, open(os.path.join(P, "bench_r02_multi_" + f[len(mtag) + 1:]), "w"), indent=1)
print("done")
