N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 2>gpurun_out/scale_n$N.err | grep '^{"metric"' > gpurun_out/scale_n$N.json; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/scale_n$N.json'));print(d['n_gpus'],round(d['value']),round(d['ms_per_step'],3),round(d['e2e']['value']),d['clocks'])"
