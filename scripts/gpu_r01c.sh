python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/r01c_bench.json 2> gpurun_out/r01c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r01c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01c_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])
for k,v in d['kernels_ms_per_step'].items(): print(k, round(v['ms_per_step'],3), v['launches_per_step'])
PY
python scripts/dbg_tc_bwd.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bwd2 -s 4 -c 1 -o gpurun_out/r01e_bwd2 python scripts/dbg_tc_bwd.py > gpurun_out/ncu5.log 2>&1
echo "ncu rc=$?"
