( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 ) | tail -3
timeout 900 python scripts/fuzz_parity.py 200 47 2>&1 | tail -3
