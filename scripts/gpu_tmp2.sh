for i in 1 2; do timeout 900 python -m pytest tests/test_gpu_simulation.py -q -m gpu 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | cut -c1-300 | head -6; done
