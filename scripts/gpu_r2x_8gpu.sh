out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > $out/r2x_bench_strong_8gpu.json 2> $out/r2x_bench_strong_8gpu.err
timeout 300 $T --master-port 29612 bench.py --gpus 8 --steps 5 --warmup 3 --scaling weak > $out/r2x_bench_weak_8gpu.json 2> $out/r2x_bench_weak_8gpu.err
timeout 600 $T --master-port 29613 bench.py --gpus 8 --steps 2 --warmup 3 --workload cfg4_20k+5k_L30_Q100_70kx128z > $out/r2x_bench_cfg4_8gpu.json 2> $out/r2x_bench_cfg4_8gpu.err
rm -f $out/r2x_cfg5_8gpu.jsonl
timeout 600 $T --master-port 29614 scripts/sweep_cfg5.py --L 15 --rows 1e6,1e7,1e8 --steps 1 --out $out/r2x_cfg5_8gpu.jsonl > $out/r2x_cfg5_8gpu_L15.log 2>&1
timeout 400 $T --master-port 29615 scripts/sweep_cfg5.py --L 30 --rows 1e6,1e7 --steps 1 --out $out/r2x_cfg5_8gpu.jsonl > $out/r2x_cfg5_8gpu_L30.log 2>&1
python - <<'PY'
import json
for f in ('r2x_bench_strong_8gpu','r2x_bench_weak_8gpu','r2x_bench_cfg4_8gpu'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, d['scaling'], 'value %.4g e2e %.4g ms %.1f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d.get('kernels_ms_per_step'))
    except Exception as e:
        print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-400:])
try:
    for l in open('gpurun_out/r2x_cfg5_8gpu.jsonl'):
        d=json.loads(l); print('cfg5 L',d['L'],'rows',d['rows'],'poses/s %.4g'%d['poses_per_s'],'ms',round(d['ms_per_call'],1))
except Exception as e: print('cfg5 ERR',e)
PY
