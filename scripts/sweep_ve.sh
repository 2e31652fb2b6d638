# tuning helper: each argument is "variant|VAR=value VAR=value"; times the 16-z slice of the bench workload
for a in "$@"; do
  v="${a%%|*}"; e="${a#*|}"; [ "$e" = "$a" ] && e=""
  lib=""; [ "$v" != default ] && lib="SXS_LIB_PATH=variants/$v/libfmftsaxs.so"
  env $lib $e python bench.py --steps 2 --warmup 3 --nrot 70000 --nz 16 --no-cpu-baseline 2>gpurun_out/sweep_err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$a', round(d['value']), {k:round(x,1) for k,x in d['kernels_ms_per_step'].items()}, d['gpu_launches'])" || tail -5 gpurun_out/sweep_err.txt
done
