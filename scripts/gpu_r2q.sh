mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -p no:cacheprovider -k "fit_kernel or scores_six" > gpurun_out/r2q_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2q_pytest.txt
grep -h "passed\|failed\|rc=" gpurun_out/r2q_pytest.txt | cut -c1-200
for v in default stash call callstash npi1call park parkstash parkcall npi3call; do
  unset SXS_LIB_PATH SXS_FIT_STASH
  case $v in default) ;; stash) export SXS_FIT_STASH=1;; callstash) export SXS_FIT_STASH=1 SXS_LIB_PATH=$PWD/variants/call/libfmftsaxs.so;; parkstash) export SXS_FIT_STASH=1 SXS_LIB_PATH=$PWD/variants/park/libfmftsaxs.so;; *) export SXS_LIB_PATH=$PWD/variants/$v/libfmftsaxs.so;; esac
  timeout 300 python bench.py --nz 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench16_$v.json 2> gpurun_out/r2q_bench16_$v.err
done
unset SXS_LIB_PATH SXS_FIT_STASH
for v in call npi3call; do
SXS_FIT_STASH=1 SXS_LIB_PATH=$PWD/variants/$v/libfmftsaxs.so python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -p no:cacheprovider -k "fit_kernel" > gpurun_out/r2q_pytest_$v.txt 2>&1; echo "rc=$?" >> gpurun_out/r2q_pytest_$v.txt
grep -h "passed\|failed\|rc=" gpurun_out/r2q_pytest_$v.txt | cut -c1-200
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2q_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernels_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
