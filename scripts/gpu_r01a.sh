set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r01a_pytest.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 20 --warmup 3 > gpurun_out/r01a_bench.json 2> gpurun_out/r01a_bench.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:cfconv_csr -s 6 -c 2 -o gpurun_out/r01a_cfconv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
tail -3 gpurun_out/r01a_pytest.log
