mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -p no:cacheprovider -k "fit_kernel or scores_six" > gpurun_out/r2s_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/r2s_pytest.txt
grep -h "passed\|failed\|rc=" gpurun_out/r2s_pytest.txt | cut -c1-200
for v in default ring3 ring6 ring8 inl; do
  unset SXS_LIB_PATH
  case $v in default) ;; *) export SXS_LIB_PATH=$PWD/variants/$v/libfmftsaxs.so;; esac
  timeout 300 python bench.py --nz 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench16_$v.json 2> gpurun_out/r2s_bench16_$v.err
done
unset SXS_LIB_PATH
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fit -s 1 -c 1 -o gpurun_out/prof_r2s_default -f python bench.py --nz 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernels_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
