mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_huffman.py tests/test_gpu_jpegs.py -x -q -m gpu 2>&1 | tail -3
run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py $1 gpu 240 5 2>&1 | tail -4 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo; }
echo "--- 128"; run 128
echo "--- 32"; run 32
PROFILE_DEVICE_OUT=1 JGPU_HUFF_WAVES=2 JGPU_HUFF_WAVE_PER_SM=4 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_huff_write --launch-skip 3 -c 1 --csv --log-file gpurun_out/r2/launches_huff_w2.csv python tools/profile_jpegs.py 24 gpu 240 1 > /dev/null 2>&1
grep -h "k_huff_write" gpurun_out/r2/launches_huff_w2.csv | awk -F'","' '{print $(NF-2), $NF}'
timeout 600 python tools/fuzz_jpegs_gpu.py 19 600 2>&1 | tail -2
