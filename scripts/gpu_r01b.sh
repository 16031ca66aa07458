python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r01b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01b_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])
for k,v in d['kernels_ms_per_step'].items(): print(k, round(v['ms_per_step'],3), v['launches_per_step'])
PY
