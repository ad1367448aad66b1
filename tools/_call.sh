run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py $1 gpu 240 5 2>&1 | tail -4 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo; }
echo "--- default 128"; run 128
echo "--- default 32"; run 32
echo "--- default 8"; run 8
echo "--- 128 host out"; timeout 300 python tools/profile_jpegs.py 128 gpu 240 3 2>&1 | tail -2 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo
echo "--- 32 host out"; timeout 300 python tools/profile_jpegs.py 32 gpu 240 3 2>&1 | tail -2 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo
timeout 900 python -m pytest tests/test_gpu_huffman.py tests/test_gpu_jpegs.py -x -q -m gpu 2>&1 | tail -3
