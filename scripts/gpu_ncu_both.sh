python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fwd2|bwd2" -s 18 -c 5 -o gpurun_out/r01g_tc2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu7.log 2>&1
echo "ncu rc=$?"
