timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
TAG=34 VARIANTS="default" WLS="4k420_b256 4kgray_b256 mixed_stress" bash tools/ab.sh
for wl in 4k420_b256 4kgray_b256; do python bench.py --workload $wl --yuv --no-e2e --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('planes $wl', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['parity']['mismatching_images'])"; done
