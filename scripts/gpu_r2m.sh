mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py -m gpu -q -rA -p no:cacheprovider -k "dense_scan" > gpurun_out/r2m_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest.txt
grep -h "passed\|failed\|rc=\|dense\|Error\|error\|assert" gpurun_out/r2m_pytest.txt | cut -c1-300 | head -30
