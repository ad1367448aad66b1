mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ncu --set full --import-source on --clock-control none -k regex:k_tk -c 1 -o gpurun_out/r2/tk_b_420_b16 -f python bench.py --workload 4k420_b16 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r2/ncu_7.log 2>&1
tail -2 gpurun_out/r2/ncu_7.log | cut -c1-200
