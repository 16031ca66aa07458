python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"fwd2|bwd2|linear_tc|prior_csr|nl_fill" -s 40 -c 12 -o gpurun_out/r01j_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
echo "full rc=$?"
