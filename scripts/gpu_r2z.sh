mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cross_dense -s 3 -c 1 -o gpurun_out/prof_r2z_dense -f python -m pytest tests/test_gpu_scale.py -m gpu -q -p no:cacheprovider -k "test_dense_scan_topk and not reference" > gpurun_out/r2z_ncu.log 2>&1
tail -2 gpurun_out/r2z_ncu.log | cut -c1-200
python scripts/correlate_e2e.py 2000000 2>&1 | tail -8
