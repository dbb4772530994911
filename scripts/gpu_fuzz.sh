timeout 1500 python scripts/fuzz_parity.py 150 23 2>&1 | tail -3
timeout 900 python scripts/fuzz_parity.py 100 31 2>&1 | tail -3
