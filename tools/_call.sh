timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decoder.py tests/test_gpu_jpegs.py tests/test_gpu_pack.py -x -q 2>&1 | tail -3
TAG=28 VARIANTS="default" WLS="4k420_b256 mixed_stress" bash tools/ab.sh
python tools/kind_bench.py 2>&1 | grep -E "1000x563|1537x771"
