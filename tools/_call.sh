timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python tools/kind_bench.py 2>&1 | grep -E "gray"
TAG=35 VARIANTS="default" WLS="mixed_stress" bash tools/ab.sh
