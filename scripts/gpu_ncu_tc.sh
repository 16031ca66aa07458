python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:filter_cfconv -s 12 -c 3 -o gpurun_out/r01c_tc_v1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu3.log
