mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TAG=22 VARIANTS="default" WLS="4k420_b256 4k422_b128 4kgray_b256 4k444_b64 4k440_b128 mixed_stress 1080p420_b512" bash tools/ab.sh
python bench.py > gpurun_out/r2/bench_22_full.json 2> gpurun_out/r2/bench_22_full.err; tail -c 3000 gpurun_out/r2/bench_22_full.json
