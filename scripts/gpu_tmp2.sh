timeout 1200 python -m pytest tests -q -m gpu 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | cut -c1-330 | head -12
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_cur.json")); print(round(d["value"]), "timestep*mol/s", round(d["ms_per_step"],3), "ms/step  e2e", round(d["e2e"]["value"]))
for k,v in d["kernels_ms_per_step"].items():
    if v["ms_per_step"]>0.05: print(" ", k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
python scripts/dbg_w16_err.py 2>&1 | tail -4
