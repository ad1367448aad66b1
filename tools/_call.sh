for n in 128 32; do
PROFILE_DEVICE_OUT=1 python tools/profile_jpegs.py $n gpu 240 4 2>&1 | tail -4
python tools/profile_jpegs.py $n gpu 240 3 2>&1 | tail -3
done
