( time timeout 900 python -m pytest tests -x -q -m gpu -k "golden or long" 2>&1 ) 2>&1 | tail -6
