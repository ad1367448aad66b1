mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r2/bench_huff3.json 2> gpurun_out/r2/bench_huff3.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2/bench_huff3.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],4),"frac",round(d["roofline"]["frac"],4),"e2e",round(d["e2e"]["value"]),"pack",round(d["e2e_pack"]["value"]))
j=d["e2e_jpeg"]
print({k:(round(v) if isinstance(v,float) else v) for k,v in j.items() if "value" in k or "threads" in k})
PY
