# usage: TAG=n VARIANTS="default tlow spin" WLS="4k420_b256 ..." bash tools/ab.sh
mkdir -p gpurun_out/r2
for v in $VARIANTS; do
  if [ "$v" = default ]; then unset JGPU_LIB_PATH; else export JGPU_LIB_PATH=$PWD/jpeg_gpu_b200/libjpeg_gpu_b200.$v.so; fi
  for wl in $WLS; do
    timeout 300 python bench.py --workload $wl --no-e2e --no-cpu --no-extra > gpurun_out/r2/bench_${TAG}_${v}_$wl.json 2>> gpurun_out/r2/bench_${TAG}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2/bench_${TAG}_${v}_$wl.json")); print("$v $wl",round(d["ms_per_step"],4),"ms frac",round(d["roofline"]["frac"],4), d["parity"]["checked"], d["parity"]["mismatching_images"])
except Exception as e: print("$v $wl","failed",e)
PY
  done
done
