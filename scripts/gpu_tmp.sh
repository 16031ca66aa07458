timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "cfconv or csr or operator or fp32" 2>&1 | tail -3
for u in 4 8; do echo "== FMD_CSR128_U=$u"; FMD_CSR128_U=$u timeout 300 python scripts/time_edge_kernels.py 2>&1 | tail -8; done | tee gpurun_out/edge_time.log
