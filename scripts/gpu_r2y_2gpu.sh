out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_scale.py -m gpu -q -rA -p no:cacheprovider -k "two_devices" > $out/r2y_pytest_2gpu.txt 2>&1; echo "rc=$?" >> $out/r2y_pytest_2gpu.txt
grep -h "passed\|failed\|rc=\|rows,\|Error" $out/r2y_pytest_2gpu.txt | cut -c1-250
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $out/r2y_bench_strong_2gpu.json 2> $out/r2y_bench_strong_2gpu.err
python - <<'PY'
import json
for f in ('r2y_bench_strong_2gpu',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, d['scaling'], 'value %.4g e2e %.4g ms %.1f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d.get('kernels_ms_per_step'))
    except Exception as e:
        print(f,'ERR',e, open('gpurun_out/%s.err'%f).read()[-600:])
PY
