timeout 900 python scripts/fuzz_parity.py 150 1 2>&1 | tail -14
timeout 900 python scripts/fuzz_parity.py 150 7 2>&1 | tail -14
