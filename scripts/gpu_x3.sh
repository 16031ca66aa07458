# BF16x3 fp32-emulation GEMM + edge kernels: parity tests, per-shape timing, fp32 bench with per-kernel times
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/x3_pytest.log
timeout 300 python scripts/time_x3.py 2>&1 | tee gpurun_out/x3_time.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/cfg2_fp32_x3.json 2> gpurun_out/cfg2f_x3.err; echo "fp32 rc=$?"; tail -3 gpurun_out/cfg2f_x3.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/cfg2_fp32_x3.json")); print(round(d["value"]), "timestep*mol/s", round(d["ms_per_step"],3), "ms/step")
for k,v in d["kernels_ms_per_step"].items(): print(" ", k, round(v["ms_per_step"],3), v["launches_per_step"])
PY
