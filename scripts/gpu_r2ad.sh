out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $out/pytest_gpu_r2ad.txt 2>&1; echo "rc=$?" >> $out/pytest_gpu_r2ad.txt
grep -h "passed\|failed\|rc=\|^FAILED\|^ERROR" $out/pytest_gpu_r2ad.txt | cut -c1-250
grep -h "rows beyond" $out/pytest_gpu_r2ad.txt | cut -c1-260 > $out/r2ad_parity_report.txt
rm -f $out/r2ad_cfg5_1gpu.jsonl
python scripts/sweep_cfg5.py --L 30,40 --rows 1e5,1e6 --steps 1 --out $out/r2ad_cfg5_1gpu.jsonl > $out/r2ad_cfg5.log 2>&1
python -c "
import json
for l in open('gpurun_out/r2ad_cfg5_1gpu.jsonl'):
    d=json.loads(l); print('cfg5 L',d['L'],'rows',d['rows'],'poses/s %.4g'%d['poses_per_s'],'ms',round(d['ms_per_call'],1),{k:round(v,1) for k,v in d['kernels_ms'].items()})
"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('cfg3', 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernels_ms_per_step'].items()})"
