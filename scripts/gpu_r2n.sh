mkdir -p gpurun_out
for v in 0 1 2 3; do
  SXS_DENSE_VARIANT=$v python -m pytest tests/test_gpu_scale.py -m gpu -q -rA -p no:cacheprovider -k "dense_scan" > gpurun_out/r2n_pytest_v$v.txt 2>&1; echo "rc=$?" >> gpurun_out/r2n_pytest_v$v.txt
  echo "variant $v"; grep -h "passed\|failed\|rc=\|dense scan of\|Error\|error" gpurun_out/r2n_pytest_v$v.txt | cut -c1-300 | head -8
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fit -s 1 -c 1 -o gpurun_out/prof_r2n_default -f python bench.py --nz 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_ncu_default.log 2>&1
tail -3 gpurun_out/r2n_ncu_default.log
