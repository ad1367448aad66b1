timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
TAG=36 VARIANTS="default" WLS="4k420_b256 4k422_b128 4k444_b64 1080p420_b512 mixed_stress" bash tools/ab.sh
KIND_MODES=420,444 KIND_SIZES=1920x1080,2048x1080,3840x2160,1537x771,960x540 python tools/kind_bench.py
