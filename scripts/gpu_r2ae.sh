SXS_LIB_PATH=$PWD/variants/carry/libfmftsaxs.so python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -p no:cacheprovider -k "fit_kernel" 2>&1 | grep -h "passed\|failed" | cut -c1-200
for v in default carry carry5 carryinl default carry; do
  unset SXS_LIB_PATH
  case $v in default) ;; *) export SXS_LIB_PATH=$PWD/variants/$v/libfmftsaxs.so;; esac
  timeout 300 python bench.py --nz 16 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$v', 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernels_ms_per_step'].items()})"
done
