mkdir -p gpurun_out/r2
python tools/pcie_probe.py > gpurun_out/r2/pcie_probe_n1.txt 2>&1; cat gpurun_out/r2/pcie_probe_n1.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py > gpurun_out/r2/pcie_probe_n2.txt 2>&1; tail -4 gpurun_out/r2/pcie_probe_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err; tail -3 gpurun_out/r2/bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2/bench_n2.json"))
print("N",d["n_gpus"],"value",round(d["value"]),"ms",round(d["ms_per_step"],4),"frac",round(d["roofline"]["frac"],4),"parity",d["parity"])
print("e2e",d["e2e"]and (round(d["e2e"]["value"]), d["e2e"]["staging"]),"pack",d["e2e_pack"] and round(d["e2e_pack"]["value"]))
print("jpeg",{k:(round(v) if isinstance(v,float) else v) for k,v in (d["e2e_jpeg"] or {}).items() if "value" in k})
for k,v in (d["extra"] or {}).items(): print(k, round(v["value"]), "frac", round(v["roofline"]["frac"],4), v["parity"]["checked"], v["parity"]["mismatching_images"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 | head -c 300
