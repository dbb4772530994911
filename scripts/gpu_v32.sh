( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 ) | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg2 value %.4g e2e %.4g cpu %.4g launches %s traffic %s frac %.4g' % (j['value'], j['e2e']['value'], j['cpu_baseline']['value'], j['gpu_launches'], j['roofline']['traffic'], j['roofline']['frac']))"
timeout 600 python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "device-resident|e2e:" | tail -2
