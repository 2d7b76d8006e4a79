timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --steps 40 > gpurun_out/final2_n1_config2.json 2> gpurun_out/final2_n1_config2.err
timeout 300 python bench.py --config 3 --no-cpu-baseline > gpurun_out/final2_n1_config3.json 2>/dev/null
python - <<'PY'
import json
for f in ["final2_n1_config2","final2_n1_config3"]:
    d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
    print(f, "%.4e"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], {k:round(v,3) for k,v in d["phases_ms"].items()}, "launches", d.get("gpu_launches"), "traffic", d["roofline"]["traffic"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
