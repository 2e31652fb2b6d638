# tuning helper: time library variants (scripts/build_variant.py) on the 16-z slice of the bench workload
for v in default "$@"; do
  lib=""; [ "$v" != default ] && lib="SXS_LIB_PATH=variants/$v/libfmftsaxs.so"
  env $lib python bench.py --steps 2 --warmup 3 --nrot 70000 --nz 16 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), {k:round(x,1) for k,x in d['kernels_ms_per_step'].items()}, d['gpu_launches'])"
done
