( timeout 900 python -m pytest tests -x -q -m gpu -k "cfg5 or long or near or cfg4-3kbp" 2>&1 ) | tail -3
export WFAGPU_TRACE=1
timeout 900 python scripts/long_reads.py 16 2>&1 | grep -E "bp:|mode=2" | tail -12
