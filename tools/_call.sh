timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_huffman.py tests/test_gpu_jpegs.py tests/test_gpu_decoder.py -x -q 2>&1 | tail -3
PROFILE_DEVICE_OUT=1 python tools/profile_jpegs.py 128 gpu 240 3 2>&1 | tail -3
TAG=32 VARIANTS="default" WLS="4k420_b256" bash tools/ab.sh
