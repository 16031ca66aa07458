set -x
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/tools/dist_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n$N.err
python -c "
import json;d=json.load(open('gpurun_out/bench_n$N.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'])"
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'])"
