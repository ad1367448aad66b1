run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py $1 gpu 240 4 2>&1 | tail -3 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo; }
echo "--- default 128"; run 128
echo "--- default 32"; run 32
echo "--- first 24 MB: 32"; JGPU_FIRST_MB=24 run 32
echo "--- first 24 MB: 128"; JGPU_FIRST_MB=24 run 128
for v in w4 w16; do export JGPU_LIB_PATH=$PWD/jpeg_gpu_b200/libjpeg_gpu_b200.$v.so; echo "--- $v 128"; run 128; done
