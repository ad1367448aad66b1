mkdir -p gpurun_out/r2
python bench.py > gpurun_out/r2/bench_8.json 2> gpurun_out/r2/bench_8.err; tail -5 gpurun_out/r2/bench_8.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2/bench_8.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],4),"frac",round(d["roofline"]["frac"],4),"parity",d["parity"])
print("e2e",d["e2e"]and round(d["e2e"]["value"]),"pack",d["e2e_pack"] and round(d["e2e_pack"]["value"]))
print("jpeg",{k:(round(v) if isinstance(v,float) else v) for k,v in (d["e2e_jpeg"] or {}).items() if "value" in k})
for k,v in (d["extra"] or {}).items(): print(k, round(v["value"]), "frac", round(v["roofline"]["frac"],4), v["parity"]["checked"], v["parity"]["mismatching_images"])
print("cpu",d["cpu_baseline"])
PY
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2/bench_8_ref.json 2>gpurun_out/r2/bench_8_ref.err; cat gpurun_out/r2/bench_8_ref.json | head -c 900; tail -3 gpurun_out/r2/bench_8_ref.err
