mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for wl in 4k420_b256 4k422_b128 4kgray_b256 4k444_b64; do
  python bench.py --workload $wl --no-e2e --no-cpu > gpurun_out/r2/bench_7_$wl.json 2> gpurun_out/r2/bench_7.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/bench_7_$wl.json")); print("$wl",round(d["ms_per_step"],4),"ms frac",round(d["roofline"]["frac"],4), d.get("parity"))
except Exception as e: print("$wl","failed",e)
PY
done
