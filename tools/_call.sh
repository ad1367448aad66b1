run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py $1 gpu 240 5 2>&1 | tail -4 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo; }
echo "--- default 128"; run 128
echo "--- default 32"; run 32
echo "--- threads 6 128"; PROFILE_THREADS=6 run 128
echo "--- threads 12 128"; PROFILE_THREADS=12 run 128
echo "--- threads 16 128"; PROFILE_THREADS=16 run 128
JGPU_TRACE=1 PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py 128 gpu 240 1 2>&1 | grep "jgpu_decode_jpegs" | tail -1
JGPU_TRACE=1 PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py 32 gpu 240 1 2>&1 | tail -9
timeout 900 python -m pytest tests/test_gpu_huffman.py tests/test_gpu_jpegs.py -x -q -m gpu 2>&1 | tail -3
