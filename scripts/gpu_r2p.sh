mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fit -s 1 -c 1 -o gpurun_out/prof_r2p_default -f python bench.py --nz 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_ncu_default.log 2>&1
tail -2 gpurun_out/r2p_ncu_default.log | cut -c1-200
