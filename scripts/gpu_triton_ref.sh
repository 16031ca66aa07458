export TRITON_CACHE_DIR=/tmp/triton_cache
for c in 0 1; do
timeout 1200 python scripts/bench_triton_reference.py --steps 100 --gptq w16a16 --compile $c > gpurun_out/triton_ref_w16_c$c.json 2> gpurun_out/triton_ref_w16_c$c.err; echo "rc=$?"; tail -2 gpurun_out/triton_ref_w16_c$c.err | cut -c1-300; cat gpurun_out/triton_ref_w16_c$c.json
done
timeout 900 python scripts/bench_triton_reference.py --steps 60 --gptq none --compile 0 > gpurun_out/triton_ref_fp32_c0.json 2> gpurun_out/triton_ref_fp32_c0.err; echo "rc=$?"; cat gpurun_out/triton_ref_fp32_c0.json
