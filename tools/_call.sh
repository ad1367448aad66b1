mkdir -p gpurun_out/r2
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -q -x -k "test_rows_of_every_alignment or test_task_counts or test_planes_out or (test_parity_by_subsampling and fused) or test_parity_mixed_batch" > gpurun_out/r2/sanitize_tk2_$tool.log 2>&1
  tail -3 gpurun_out/r2/sanitize_tk2_$tool.log
done
