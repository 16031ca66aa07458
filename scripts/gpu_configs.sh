python bench.py --steps 10 --warmup 3 --no-cpu-baseline --n-beads 500 --batch 64 --blocks 5 > gpurun_out/cfg5.json 2> gpurun_out/cfg5.err; echo "cfg5 rc=$?"; tail -2 gpurun_out/cfg5.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 1024 > gpurun_out/cfg3_1gpu.json 2> gpurun_out/cfg3.err; echo "cfg3 rc=$?"; tail -2 gpurun_out/cfg3.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --n-beads 54 --batch 128 > gpurun_out/cfg2_literal54.json 2> gpurun_out/cfg2l.err; echo "cfg2-54 rc=$?"; tail -2 gpurun_out/cfg2l.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/cfg2_fp32.json 2> gpurun_out/cfg2f.err; echo "fp32 rc=$?"; tail -2 gpurun_out/cfg2f.err
python - <<'PY'
import json
for f in ["cfg5","cfg3_1gpu","cfg2_literal54","cfg2_fp32"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), "timestep*mol/s", round(d["ms_per_step"],3), "ms/step edges", d["edges"], "nodes", d["nodes"])
    except Exception as e: print(f, "ERR", e)
PY
