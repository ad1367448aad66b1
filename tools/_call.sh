mkdir -p gpurun_out/r2
PROFILE_DEVICE_OUT=1 JGPU_HUFF_WAVES=2 JGPU_HUFF_WAVE_PER_SM=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_huff_write --launch-skip 3 --launch-count 1 -o /tmp/write_f python tools/profile_jpegs.py 24 gpu 240 1 > /dev/null 2>&1
ncu -i /tmp/write_f.ncu-rep --page raw --csv > gpurun_out/r2/ncu_write_final2_raw.csv
PROFILE_DEVICE_OUT=1 JGPU_HUFF_WAVES=2 JGPU_HUFF_WAVE_PER_SM=4 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_ --launch-skip 21 -c 9 --csv --log-file gpurun_out/r2/launches_jpeg_final_9files.csv python tools/profile_jpegs.py 24 gpu 240 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2/launches_jpeg_final_9files.csv")) if len(r)>10]
h=rows[0]
acc={}
for r in rows[1:]:
    d=dict(zip(h,r))
    acc.setdefault(d["ID"],[d["Kernel Name"][:60], d["Grid Size"]]).append(d["Metric Value"])
for k,v in acc.items(): print(k, v)
PY
