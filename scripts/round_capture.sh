#!/bin/bash
# One GPU-box session: GPU tests, full-size bench (both arms), ncu launch list, DRAM traffic of every kernel at full
# size, full-set captures of K4 and K3 at the 16-z size.  Outputs under gpurun_out/ with the tag given as $1.
tag=${1:-r1c}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $out/pytest_gpu_$tag.txt
cat $out/pytest_gpu_$tag.txt
python bench.py --steps 5 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err
tail -c 600 $out/bench_$tag.json
( time python bench.py --impl reference --steps 2 --warmup 1 ) > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
tail -c 400 $out/bench_ref_$tag.json; tail -4 $out/bench_ref_$tag.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file $out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_fit|k_cross|k_translate_rt" --launch-skip 9 --launch-count 3 \
    -o $out/prof_${tag}_k34 -f python bench.py --steps 1 --warmup 3 --nrot 70000 --nz 16 --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
tail -2 $out/ncu_full_$tag.log
