mkdir -p gpurun_out/r2
python bench.py > gpurun_out/r2/bench_final2.json 2> gpurun_out/r2/bench_final2.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/r2/launches_final2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > /dev/null 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/r2/bench_final2.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],4),"frac",round(d["roofline"]["frac"],4),"e2e",round(d["e2e"]["value"]),"pack",round(d["e2e_pack"]["value"]),"jpeg",round(d["e2e_jpeg"]["value"]),round(d["e2e_jpeg"]["device_out_value"]),round(d["e2e_jpeg"]["yuv_out_value"]), "cpu", round(d["cpu_baseline"]["value"]))
for k,v in d["extra"].items(): print(k, round(v["value"]), round(v["ms_per_step"],4), round(v["roofline"]["frac"],4), v["parity"]["mismatching_images"])
PY
