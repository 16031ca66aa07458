# round-2 evidence of the W16A16 step: bench line, ncu launch list of the same command, --set full of the top kernels
T=${1:-r02h}
timeout 1200 python bench.py --steps 50 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-triton-baseline > gpurun_out/ncu_l.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"filter_cfconv|linear_chain|prior_csr|nl_|edge_grad" -s 40 -c 24 -o gpurun_out/${T}_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-triton-baseline > gpurun_out/ncu_f.log 2>&1; echo "full rc=$?"
