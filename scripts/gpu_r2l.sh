mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -p no:cacheprovider -k "fit_kernel or scores_six" > gpurun_out/r2l_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest.txt
for v in default nobatch half ring6 ring3; do
  if [ $v = default ]; then unset SXS_LIB_PATH; else export SXS_LIB_PATH=$PWD/variants/$v/libfmftsaxs.so; fi
  timeout 300 python bench.py --nz 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench16_$v.json 2> gpurun_out/r2l_bench16_$v.err
done
unset SXS_LIB_PATH
grep -h "passed\|failed\|rc=" gpurun_out/r2l_pytest.txt | cut -c1-200
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernels_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
