# tuning helper: time the 16-z slice of the bench workload under different environment settings
#   scripts/sweep_env.sh "SXS_CROSS_GROUP=1" "SXS_CROSS_GROUP=2" ...      (each argument: space-separated VAR=value list)
for v in "$@"; do
  env $v python bench.py --steps 2 --warmup 3 --nrot 70000 --nz 16 --no-cpu-baseline 2>gpurun_out/sweep_err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), {k:round(x,1) for k,x in d['kernels_ms_per_step'].items()}, d['gpu_launches'])" || tail -5 gpurun_out/sweep_err.txt
done
