run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py $1 gpu 240 4 2>&1 | tail -3 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo; }
for v in default loop; do
  if [ "$v" = default ]; then unset JGPU_LIB_PATH; else export JGPU_LIB_PATH=$PWD/jpeg_gpu_b200/libjpeg_gpu_b200.$v.so; fi
  echo "--- $v 128"; run 128
  echo "--- $v q98 noise 60, 32 files"; PROFILE_NOISE=60 PROFILE_Q=98 run 32
  echo "--- $v q95, 64 files"; PROFILE_Q=95 run 64
done
unset JGPU_LIB_PATH
timeout 600 python -m pytest tests/test_gpu_huffman.py -x -q -m gpu 2>&1 | tail -2
