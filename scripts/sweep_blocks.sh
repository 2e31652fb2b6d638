for b in 1 2 3; do for v in default ldcs; do
  lib=""; [ "$v" != default ] && lib="SXS_LIB_PATH=variants/$v/libfmftsaxs.so"
  env $lib SXS_FIT_BLOCKS_PER_SM=$b python bench.py --steps 2 --warmup 3 --nrot 70000 --nz 16 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v blocks/SM=$b', round(d['value']), {k:round(x,1) for k,x in d['kernels_ms_per_step'].items()})"
done; done
