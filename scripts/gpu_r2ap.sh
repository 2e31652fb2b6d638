SXS_LIB_PATH=$PWD/variants/t288/libfmftsaxs.so python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fit_kernel" 2>&1 | tail -1
for v in default t288 t320 default t288; do
  unset SXS_LIB_PATH
  case $v in default) ;; *) export SXS_LIB_PATH=$PWD/variants/$v/libfmftsaxs.so;; esac
  timeout 300 python bench.py --nz 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2ap_$v.json 2> gpurun_out/r2ap_$v.err
  python -c "
import sys,json
try:
    d=json.loads(open('gpurun_out/r2ap_$v.json').read().strip().split('\n')[-1]); print('$v', 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernels_ms_per_step'].items()})
except Exception as e: print('$v', 'ERR', open('gpurun_out/r2ap_$v.err').read()[-200:])"
done
