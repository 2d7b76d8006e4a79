# final single-GPU evidence run: tests, default bench (both arms), config 3/4 lines, ncu launch list
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/final_n1_config2.json 2> gpurun_out/final_n1_config2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference_arm.json 2> gpurun_out/final_reference_arm.err
timeout 300 python bench.py --config 3 --no-cpu-baseline > gpurun_out/final_n1_config3.json 2>/dev/null
timeout 300 python bench.py --config 4 --no-cpu-baseline --steps 5 > gpurun_out/final_n1_config4.json 2>/dev/null
timeout 300 python bench.py --config 4 --limited --no-cpu-baseline --steps 5 > gpurun_out/final_n1_config4_limited.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_config2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_b.log 2>&1
python - <<'PY'
import json
for f in ["final_n1_config2","final_reference_arm","final_n1_config3","final_n1_config4","final_n1_config4_limited"]:
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, "%.4e"%d["value"], "ms %.3f"%d.get("ms_per_step",0), "e2e %.3e"%d["e2e"]["value"], {k:round(v,3) for k,v in d.get("phases_ms",{}).items()}, d.get("clocks"), d.get("roofline"), d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
