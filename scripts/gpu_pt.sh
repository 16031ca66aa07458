N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python scripts/bench_pt.py --steps ${STEPS:-400} 2> gpurun_out/pt_n1.err | grep '^{"metric"' | tee gpurun_out/pt_n1.json; tail -3 gpurun_out/pt_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/bench_pt.py --steps ${STEPS:-400} 2> gpurun_out/pt_n$N.err | grep '^{"metric"' | tee gpurun_out/pt_n$N.json; tail -3 gpurun_out/pt_n$N.err
fi
