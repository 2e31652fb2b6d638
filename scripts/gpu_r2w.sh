mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py -m gpu -q -rA -p no:cacheprovider -k "ft_rows_to_indices or dense_scan" > gpurun_out/r2w_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2w_pytest.txt
grep -h "passed\|failed\|rc=\|ft rows ->\|dense scan of\|Error\|assert\|top-48" gpurun_out/r2w_pytest.txt | cut -c1-400 | head -20
for v in 1 2; do SXS_DENSE_VARIANT=$v python -m pytest tests/test_gpu_scale.py -m gpu -q -rA -p no:cacheprovider -k "test_dense_scan_topk and not reference" 2>&1 | grep "dense scan of" | head -2; done
