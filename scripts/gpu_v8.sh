( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 ) | tail -8
export WFAGPU_TRACE=1
timeout 900 python scripts/long_reads.py 16 2>&1 | grep -v "tier [0-2] " | tail -14
