mkdir -p gpurun_out/r2
JGPU_LIB_PATH=$PWD/jpeg_gpu_b200/libjpeg_gpu_b200.trace.so python tools/mcu_trace.py 4k420_b256 2>&1 | head -4
for lib in "" w2us; do
for wl in 4k420_b256 4k422_b128 4kgray_b256 4k444_b64 mixed_stress; do
  if [ -n "$lib" ]; then export JGPU_LIB_PATH=$PWD/jpeg_gpu_b200/libjpeg_gpu_b200.$lib.so; else unset JGPU_LIB_PATH; fi
  python bench.py --workload $wl --no-e2e --no-cpu > gpurun_out/r2/bench_3_mcu${lib}_$wl.json 2> gpurun_out/r2/bench_3.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/bench_3_mcu${lib}_$wl.json")); print("$wl","mcu$lib",round(d["ms_per_step"],4),"ms frac",round(d["roofline"]["frac"],4), d.get("parity"))
except Exception as e: print("$wl","failed",e)
PY
done
done
unset JGPU_LIB_PATH
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
