python scripts/dbg_fp32_err.py 2>&1 | tail -14
