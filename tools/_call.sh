run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py $1 gpu 240 5 2>&1 | tail -4 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo; }
echo "--- default 128"; run 128
export JGPU_LIB_PATH=$PWD/jpeg_gpu_b200/libjpeg_gpu_b200.c128.so
echo "--- c128 wave 12/SM: 128"; JGPU_HUFF_WAVE_PER_SM=12 run 128
echo "--- c128 wave 24/SM: 128"; JGPU_HUFF_WAVE_PER_SM=24 run 128
echo "--- c128 wave 12/SM: 32"; JGPU_HUFF_WAVE_PER_SM=12 run 32
JGPU_HUFF_WAVE_PER_SM=12 PROFILE_DEVICE_OUT=1 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_huff --launch-skip 18 -c 6 --csv --log-file gpurun_out/r2/launches_huff_c128.csv python tools/profile_jpegs.py 24 gpu 240 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2/launches_huff_c128.csv")) if len(r)>10]
h=rows[0]
acc={}
for r in rows[1:]:
    d=dict(zip(h,r))
    acc.setdefault(d["ID"],[d["Kernel Name"][15:45], d["Grid Size"]]).append(d["Metric Value"])
for k,v in acc.items(): print(k, v)
PY
