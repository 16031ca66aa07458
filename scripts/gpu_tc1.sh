timeout 300 python scripts/dbg_tc_fwd.py 2>&1 | tail -20
