mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2/bench_final3.json 2> gpurun_out/r2/bench_final3.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2/bench_final3.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],4),"frac",round(d["roofline"]["frac"],4),"e2e",round(d["e2e"]["value"]),"pack",round(d["e2e_pack"]["value"]), "launches", d["gpu_launches"], "parity", d["parity"]["mismatching_images"])
j=d["e2e_jpeg"]
print({k:(round(v) if isinstance(v,float) else v) for k,v in j.items() if "value" in k or "threads" in k})
for k,v in d["extra"].items(): print(k, round(v["value"]), round(v["roofline"]["frac"],4), v["parity"]["mismatching_images"])
PY
python tools/make_jpeg.py 3840 2160 /tmp/f4k.jpg 2>/dev/null || python - <<PY
import io, numpy as np
from PIL import Image
w,h=3840,2160
rng=np.random.default_rng(1); yy,xx=np.mgrid[0:h,0:w]
base=np.stack([(xx*5+yy*3)%256,(yy*7+xx)%256,(xx*2+yy*9)%256],-1)
pic=np.clip(base+rng.integers(-24,25,size=base.shape),0,255).astype(np.uint8)
Image.fromarray(pic).save("/tmp/f4k.jpg","JPEG",quality=85,subsampling=2)
PY
ls -la /tmp/f4k.jpg
./jpeg_gpu_b200/jpeg_gpu_cli -n 50 -e gpu /tmp/f4k.jpg 2>&1 | tail -3
./jpeg_gpu_b200/jpeg_gpu_cli -n 50 -e gpu -o yuv /tmp/f4k.jpg 2>&1 | tail -2
