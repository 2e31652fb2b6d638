out=gpurun_out; tag=r2al; mkdir -p $out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $out/pytest_gpu_$tag.txt 2>&1; echo "rc=$?" >> $out/pytest_gpu_$tag.txt
grep -h "passed\|failed\|rc=\|^FAILED\|^ERROR" $out/pytest_gpu_$tag.txt | cut -c1-250
python bench.py --steps 5 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err
tail -c 300 $out/bench_$tag.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file $out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_launches_$tag.log 2>&1
python - <<'PY'
import json
for f in ('bench_r2al','bench_ref_r2al'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, 'value %.4g e2e %.4g ms %.1f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d.get('kernels_ms_per_step'), d.get('parity'))
    except Exception as e:
        print(f,'ERR',e)
PY
