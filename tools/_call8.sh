mkdir -p gpurun_out/r2
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/bench_n8_final.json 2> gpurun_out/r2/bench_n8_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2/bench_n8_final.json"))
print("n8 value",round(d["value"]),"ms",round(d["ms_per_step"],4),"frac",round(d["roofline"]["frac"],4),"parity",d["parity"]["mismatching_images"],"e2e",round(d["e2e"]["value"]),"pack",round(d["e2e_pack"]["value"]),"jpeg",round(d["e2e_jpeg"]["value"]),round(d["e2e_jpeg"]["device_out_value"]))
for k,v in d["extra"].items(): print(k, round(v["value"]), round(v["roofline"]["frac"],4), v["parity"]["mismatching_images"])
PY
