python scripts/dbg_linear_tc.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 12 -c 1 -o gpurun_out/r01f_lintc python scripts/dbg_linear_tc.py > gpurun_out/ncu6.log 2>&1
echo "ncu rc=$?"
