# ncu on k_fit (lean, exact objective): default 256x1 and t128x4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fit -s 1 -c 1 -o gpurun_out/prof_r2c_default -f python bench.py --nz 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_default.log 2>&1
SXS_LIB_PATH=$PWD/variants/t128x4/libfmftsaxs.so ncu --set full --clock-control none --import-source on -k regex:k_fit -s 1 -c 1 -o gpurun_out/prof_r2c_t128x4 -f python bench.py --nz 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_t128x4.log 2>&1
ls -la gpurun_out/*.ncu-rep
