"""Attribute the per-SASS-instruction counts of an ncu report to CUDA source lines.
ncu's CSV source page has no line column, so the kernel is disassembled with line info
(cuobjdump -xelf + nvdisasm -g) from the SAME .so and joined by instruction order.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel substring> [libwva_b200.so] [top N]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "inferno-autoscaler_b200", "libwva_b200.so")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel with their (file, line)
lines, inside, cur = [], False, ("?", 0)
for l in dis:
    if l.startswith("//---") and ".text." in l:
        inside = kname in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append((cur, l.split("*/", 1)[1].strip()))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + (sys.argv[5] if len(sys.argv) > 5 else kname)], capture_output=True, text=True).stdout.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; ie, smp = h.index("Instructions Executed"), h.index("# Samples")
sass, seen = [], set()
for r in rows[hi + 1:]:
    if len(r) > ie and r[0].startswith('0x') and r[0] not in seen:
        seen.add(r[0]); sass.append(r)
assert abs(len(sass) - len(lines)) <= 8, (len(sass), len(lines))
agg, sagg = collections.Counter(), collections.Counter()
tot = ts = 0
for (loc, txt), r in zip(lines, sass):
    n, s = int(r[ie] or 0), int(r[smp] or 0)
    agg[loc] += n; sagg[loc] += s; tot += n; ts += s
print("instructions", tot, "samples", ts)
srcs = {}
for (f, ln), n in agg.most_common(top):
    if f not in srcs:
        p = [os.path.join(d, f) for d in (os.path.join(os.path.dirname(so), "csrc"),) if os.path.exists(os.path.join(d, f))]
        srcs[f] = open(p[0]).read().splitlines() if p else []
    text = srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100.0 * n / tot, 100.0 * sagg[(f, ln)] / max(ts, 1), f, ln, text))
