"""Extract the key metrics of every kernel in an .ncu-rep (ncu -i ... --page raw --csv) into a JSON list."""
import csv, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_active.max",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
out = []
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")], "report": rep.split("/")[-1]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (vals[i] + " " + units[i]).strip()
        out.append(d)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(len(out), "kernels ->", sys.argv[1])
