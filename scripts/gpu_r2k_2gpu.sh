out=gpurun_out; mkdir -p $out
rm -f $out/r2k_cfg5_2gpu.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/sweep_cfg5.py --L 15,30 --rows 1e5,1e6 --steps 1 --out $out/r2k_cfg5_2gpu.jsonl > $out/r2k_cfg5_2gpu.log 2>&1
tail -3 $out/r2k_cfg5_2gpu.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 2 --warmup 3 --workload cfg4_20k+5k_L30_Q100_70kx128z --nz 8 --nrot 20000 > $out/r2k_bench_cfg4_small_2gpu.json 2> $out/r2k_bench_cfg4_small_2gpu.err
tail -c 1500 $out/r2k_bench_cfg4_small_2gpu.json; tail -3 $out/r2k_bench_cfg4_small_2gpu.err
