mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_huffman.py tests/test_gpu_jpegs.py -x -q -m gpu 2>&1 | tail -3
run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py $1 gpu $2 5 2>&1 | tail -4 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo; }
echo "--- 128"; run 128 240
echo "--- 32"; run 32 240
echo "--- 128 no restart markers"; run 128 0
PROFILE_DEVICE_OUT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_huff_scan -c 6 --csv --log-file gpurun_out/r2/launches_huff_scan.csv python tools/profile_jpegs.py 24 gpu 240 1 > /dev/null 2>&1
grep -h "k_huff_scan" gpurun_out/r2/launches_huff_scan.csv | awk -F'","' '{print $(NF-6), $NF}'
timeout 600 python tools/fuzz_jpegs_gpu.py 23 600 2>&1 | tail -2
