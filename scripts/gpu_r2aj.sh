mkdir -p gpurun_out
python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "translation or scores_six or cfg4 or L30 or config5 or real_list or cfg3" 2>&1 | tail -3
rm -f gpurun_out/r2aj_cfg5_1gpu.jsonl
python scripts/sweep_cfg5.py --L 30,40 --rows 1e5 --steps 1 --out gpurun_out/r2aj_cfg5_1gpu.jsonl > gpurun_out/r2aj_cfg5.log 2>&1
python -c "
import json
for l in open('gpurun_out/r2aj_cfg5_1gpu.jsonl'):
    d=json.loads(l); print('cfg5 L',d['L'],'rows',d['rows'],'ms',round(d['ms_per_call'],1),{k:round(v,1) for k,v in d['kernels_ms'].items()})
"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('cfg3', 'ms %.1f'%d['ms_per_step'], {k:round(v,1) for k,v in d['kernels_ms_per_step'].items()})"
