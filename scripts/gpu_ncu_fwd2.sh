python scripts/dbg_tc_fwd.py big > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwd2 -s 2 -c 1 -o gpurun_out/r01d_fwd2 python scripts/dbg_tc_fwd.py big > gpurun_out/ncu4.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu4.log
