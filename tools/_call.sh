mkdir -p gpurun_out/r2
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "test_parity_by_subsampling or test_planes_out or test_parity_mixed_batch or test_sixteen_bit_quant" > gpurun_out/r2/sanitize_tk_racecheck2.log 2>&1
tail -3 gpurun_out/r2/sanitize_tk_racecheck2.log
TAG=23 VARIANTS="default arr1" WLS="4k420_b256 4k444_b64 4kgray_b256" bash tools/ab.sh
