# round-2 multi-GPU evidence on N GPUs of one box: bench.py (plain Langevin, weak scaling), bench.py --pt (cfg4, NCCL exchange),
# the NCCL sharded-exchange check
N=${1:-4}; T=${2:-r02k}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench rc=$?"
timeout 900 $TR bench.py --gpus $N --pt > gpurun_out/${T}_pt_cfg4_n$N.json 2> gpurun_out/${T}_pt_n$N.err; echo "pt rc=$?"
timeout 600 $TR tests/tools/dist_check.py > gpurun_out/${T}_dist_check_n$N.txt 2>&1; echo "dist_check rc=$?"; tail -3 gpurun_out/${T}_dist_check_n$N.txt
