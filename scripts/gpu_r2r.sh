mkdir -p gpurun_out
for v in npi1call npi1callstash npi1inl npi1inlstash; do
  unset SXS_LIB_PATH SXS_FIT_STASH
  case $v in npi1callstash) export SXS_FIT_STASH=1 SXS_LIB_PATH=$PWD/variants/npi1call/libfmftsaxs.so;; npi1inlstash) export SXS_FIT_STASH=1 SXS_LIB_PATH=$PWD/variants/npi1inl/libfmftsaxs.so;; *) export SXS_LIB_PATH=$PWD/variants/$v/libfmftsaxs.so;; esac
  timeout 300 python bench.py --nz 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench16_$v.json 2> gpurun_out/r2r_bench16_$v.err
done
export SXS_FIT_STASH=1 SXS_LIB_PATH=$PWD/variants/npi1call/libfmftsaxs.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fit -s 1 -c 1 -o gpurun_out/prof_r2r_npi1callstash -f python bench.py --nz 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_ncu.log 2>&1
unset SXS_LIB_PATH SXS_FIT_STASH
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernels_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
