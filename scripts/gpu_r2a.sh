set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/r2a_pytest_fused.txt 2>&1; echo "fused rc=$?" >> gpurun_out/r2a_pytest_fused.txt
SXS_LIB_PATH=$PWD/variants/exact/libfmftsaxs.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -rA -p no:cacheprovider -k "fit_kernel or scores_ or live_reference or real_list or config4 or bench_workload or device_exp" > gpurun_out/r2a_pytest_exact.txt 2>&1; echo "exact rc=$?" >> gpurun_out/r2a_pytest_exact.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
SXS_LIB_PATH=$PWD/variants/exact/libfmftsaxs.so python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_exact.json 2> gpurun_out/r2a_bench_exact.err
tail -3 gpurun_out/r2a_pytest_fused.txt gpurun_out/r2a_pytest_exact.txt
