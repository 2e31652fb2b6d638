out=gpurun_out; mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_active
for pt in "10 1e6" "15 1e6" "20 1e6" "30 1e6" "40 1e5" "15 1e7"; do
  set -- $pt
  timeout 600 ncu --metrics $M --clock-control none -k regex:"k_fit|k_cross|k_translate_rt|k_tmatrix" --csv --log-file $out/r2ak_ncu_L$1_$2.csv \
     python scripts/sweep_cfg5.py --L $1 --rows $2 --steps 1 --out $out/r2ak_tmp.jsonl > $out/r2ak_L$1_$2.log 2>&1
  echo "L $1 rows $2 rc=$?"
done
