# tuning helper: time bench variants (alternative builds of the library selected through SXS_LIB_PATH)
for v in variants_bs128.so variants_bs256.so variants_bs512.so variants_bs768.so; do
  SXS_LIB_PATH=$PWD/$v python bench.py --steps 2 --warmup 3 --nrot 70000 --nz 16 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), {k:round(x,1) for k,x in d['kernels_ms_per_step'].items()})"
done
