mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -p no:cacheprovider -k "fit_kernel or scores_six" > gpurun_out/r2aa_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2aa_pytest.txt
grep -h "passed\|failed\|rc=" gpurun_out/r2aa_pytest.txt | cut -c1-200
for v in default inl inlr3 default inl; do
  unset SXS_LIB_PATH
  case $v in default) ;; *) export SXS_LIB_PATH=$PWD/variants/$v/libfmftsaxs.so;; esac
  timeout 300 python bench.py --nz 16 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$v', 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernels_ms_per_step'].items()})"
done
