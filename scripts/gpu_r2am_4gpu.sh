out=gpurun_out; mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29911 bench.py --gpus 4 --steps 5 --warmup 3 > $out/r2am_bench_strong_4gpu.json 2> $out/r2am_bench_strong_4gpu.err
python - <<'PY'
import json
for f in ('r2am_bench_strong_4gpu',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, d['scaling'], 'value %.4g e2e %.4g ms %.1f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d.get('kernels_ms_per_step'), d['resident_equals_host_path'])
    except Exception as e:
        print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-400:])
PY
