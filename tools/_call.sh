mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pack.py tests/test_gpu_decoder.py -x -q 2>&1 | tail -5
for wl in 4k420_b256 4k422_b128 4kgray_b256 4k444_b64 4k440_b128 mixed_stress; do
  timeout 300 python bench.py --workload $wl --no-e2e --no-cpu --no-extra > gpurun_out/r2/bench_${TAG}_$wl.json 2> gpurun_out/r2/bench_${TAG}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/bench_${TAG}_$wl.json")); print("$wl",round(d["ms_per_step"],4),"ms frac",round(d["roofline"]["frac"],4), d["parity"]["checked"], d["parity"]["mismatching_images"])
except Exception as e: print("$wl","failed",e)
PY
done
