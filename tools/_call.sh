mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_huffman.py -x -q -m gpu 2>&1 | tail -3
run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py $1 gpu 240 4 2>&1 | tail -3 | awk '{print $(NF-4), $(NF-3)}' | tr '\n' ' '; echo; }
echo "--- rounds 128 files"; run 128
echo "--- rounds 32 files"; run 32
echo "--- cta 128 files"; JGPU_HUFF_SYNC=cta run 128
PROFILE_DEVICE_OUT=1 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_huff_round --launch-skip 52 -c 26 --csv --log-file gpurun_out/r2/launches_huff_rounds.csv python tools/profile_jpegs.py 16 gpu 240 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2/launches_huff_rounds.csv")) if len(r)>10]
h=rows[0]
acc={}
for r in rows[1:]:
    d=dict(zip(h,r))
    acc.setdefault(d["ID"],[d["Kernel Name"][25:60], d["Grid Size"]]).append(d["Metric Value"])
tot=0
for k,v in acc.items():
    print(k, v); tot+=float(v[2].replace(",",""))
print("total ns", tot)
PY
timeout 600 python tools/fuzz_jpegs_gpu.py 13 400 2>&1 | tail -2
