N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02s_bench_n$N.json 2> gpurun_out/r02s_bench_n$N.err; echo "bench rc=$?"
