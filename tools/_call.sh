mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_huffman.py -x -q -m gpu 2>&1 | tail -15
