timeout 1500 python scripts/fuzz_parity.py 120 11 2>&1 | tail -8
