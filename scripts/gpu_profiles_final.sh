# final-state evidence of the W16A16 step: bench line (with cpu_baseline), ncu launch list, --set full of the top kernels,
# then the other BASELINE configs on one GPU
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r01s_bench.json 2> gpurun_out/r01s_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r01s_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01s_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fwd2|bwd2|linear_chain|prior_csr|nl_kernel" -s 30 -c 14 -o gpurun_out/r01s_top python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1; echo "full rc=$?"
bash scripts/gpu_configs.sh
