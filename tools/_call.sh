set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/gputest_1.log
python bench.py > gpurun_out/r2/bench_1.json 2> gpurun_out/r2/bench_1.err
tail -c 600 gpurun_out/r2/bench_1.err
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 6 -c 1 -o gpurun_out/r2/fused_v9_b16 python bench.py --workload 4k420_b16 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2/ncu_1.log 2>&1
tail -5 gpurun_out/r2/ncu_1.log
cat gpurun_out/r2/gputest_1.log
cat gpurun_out/r2/bench_1.json | head -c 3000
