mkdir -p gpurun_out/r2
PROFILE_DEVICE_OUT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_huff_sync --launch-skip 6 --launch-count 1 -o /tmp/sync_gw python tools/profile_jpegs.py 16 gpu 240 1 > /dev/null 2>&1
PROFILE_DEVICE_OUT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_huff_write --launch-skip 2 --launch-count 1 -o /tmp/write_st python tools/profile_jpegs.py 16 gpu 240 1 > /dev/null 2>&1
ncu -i /tmp/sync_gw.ncu-rep --page raw --csv > gpurun_out/r2/ncu_sync_gw_raw.csv
ncu -i /tmp/write_st.ncu-rep --page raw --csv > gpurun_out/r2/ncu_write_staged_raw.csv
ncu -i /tmp/sync_gw.ncu-rep --page source --csv > gpurun_out/r2/ncu_sync_gw_source.csv 2>/dev/null
ncu -i /tmp/write_st.ncu-rep --page source --csv > gpurun_out/r2/ncu_write_staged_source.csv 2>/dev/null
ls -la gpurun_out/r2/ncu_*
