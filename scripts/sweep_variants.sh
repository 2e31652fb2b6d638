# tuning helper: time bench variants selected by environment (edit the list; results go to profiles/r1_k4_notes.md)
for v in "SXS_NOP=1"; do
  env $v python bench.py --steps 2 --warmup 3 --nrot 70000 --nz 16 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('${v:-default}', round(d['value']), {k:round(x,1) for k,x in d['kernels_ms_per_step'].items()}, d['gpu_launches'])"
done
