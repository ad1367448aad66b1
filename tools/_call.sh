nproc
run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py $1 gpu 240 4 2>&1 | tail -3 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo; }
for t in 0 4 8 16 32 64; do echo "--- threads $t: 128 files"; PROFILE_THREADS=$t run 128; done
for t in 0 8 16; do echo "--- threads $t: 32 files"; PROFILE_THREADS=$t run 32; done
PROFILE_THREADS=16 JGPU_TRACE=1 PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py 128 gpu 240 1 2>&1 | tail -19
