mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_translate_rt -s 2 -c 1 -o gpurun_out/prof_r2ai_translate_L30 -f python scripts/sweep_cfg5.py --L 30 --rows 1e5 --steps 1 --out gpurun_out/r2ai_tmp.jsonl > gpurun_out/r2ai_ncu.log 2>&1
tail -2 gpurun_out/r2ai_ncu.log | cut -c1-200
