mkdir -p gpurun_out/r2
nvidia-smi topo -m > gpurun_out/r2/topo_n8.txt 2>&1
nproc >> gpurun_out/r2/topo_n8.txt
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py > gpurun_out/r2/pcie_probe_n$n.txt 2> gpurun_out/r2/pcie_probe_n$n.err
  tail -4 gpurun_out/r2/pcie_probe_n$n.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/bench_n8.json 2> gpurun_out/r2/bench_n8.err
tail -c 600 gpurun_out/r2/bench_n8.json
