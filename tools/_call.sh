# what a round's last GPU call runs: every GPU test, the smoke entry, the default bench line
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2/bench_last.json 2> gpurun_out/r2/bench_last.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2/bench_last.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],4),"frac",round(d["roofline"]["frac"],4),"e2e",round(d["e2e"]["value"]),"pack",round(d["e2e_pack"]["value"]), "launches", d["gpu_launches"], "parity", d["parity"]["mismatching_images"])
j=d["e2e_jpeg"]
print({k:(round(v) if isinstance(v,float) else v) for k,v in j.items() if "value" in k or "threads" in k})
for k,v in d["extra"].items(): print(k, round(v["value"]), round(v["roofline"]["frac"],4), v["parity"]["mismatching_images"])
PY
