# the other BASELINE configs on one GPU (each a bench.py line), PT config 4 on one GPU
T=${1:-r02k}
B="--no-cpu-baseline --no-triton-baseline"
python bench.py --steps 10 --warmup 3 $B --n-beads 500 --batch 64 --blocks 5 > gpurun_out/cfg5.json 2> gpurun_out/cfg5.err; echo "cfg5 rc=$?"
python bench.py --steps 10 --warmup 3 $B --batch 1024 > gpurun_out/cfg3_1gpu.json 2> gpurun_out/cfg3.err; echo "cfg3 rc=$?"
python bench.py --steps 20 --warmup 3 $B --n-beads 54 --batch 128 > gpurun_out/cfg2_literal54.json 2> gpurun_out/cfg2l.err; echo "cfg2-54 rc=$?"
python bench.py --steps 5 --warmup 3 $B --precision fp32 > gpurun_out/cfg2_fp32.json 2> gpurun_out/cfg2f.err; echo "fp32 rc=$?"
python bench.py --pt > gpurun_out/${T}_pt_cfg4_n1.json 2> gpurun_out/${T}_pt_n1.err; echo "pt rc=$?"
python - <<PY
import json
out = {}
for f in ["cfg5", "cfg3_1gpu", "cfg2_literal54", "cfg2_fp32"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        out[f] = {k: d[k] for k in ("value", "unit", "ms_per_step", "edges_start", "edges", "nodes", "launches_per_step", "config", "kernels_ms_per_step") if k in d}
        print(f, round(d["value"]), "timestep*mol/s", round(d["ms_per_step"], 3), "ms/step edges", d["edges"], "nodes", d["nodes"])
    except Exception as e:
        print(f, "ERR", e)
json.dump(out, open("gpurun_out/${T}_other_configs.json", "w"), indent=1)
PY
