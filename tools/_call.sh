mkdir -p gpurun_out/r2 /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2/bench_final.json 2> gpurun_out/r2/bench_final.err; tail -c 400 gpurun_out/r2/bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/bench_final_ref.json 2> gpurun_out/r2/bench_final_ref.err; head -c 600 gpurun_out/r2/bench_final_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/r2/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_tk -c 1 -o /tmp/ncu/final_b256 -f python bench.py --workload 4k420_b256 --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r2/ncu_final.log 2>&1
ncu -i /tmp/ncu/final_b256.ncu-rep --page details --csv > gpurun_out/r2/final_b256_details.csv 2>/dev/null
ncu -i /tmp/ncu/final_b256.ncu-rep --page raw --csv > gpurun_out/r2/final_b256_raw.csv 2>/dev/null
