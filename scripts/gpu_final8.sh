mkdir -p gpurun_out/final
python bench.py > gpurun_out/final/bench_cfg2.json 2> gpurun_out/final/bench_cfg2.err
python -c "
import json
j=json.loads([l for l in open('gpurun_out/final/bench_cfg2.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g cpu %.4g launches %s' % (j['value'], j['e2e']['value'], j['cpu_baseline']['value'], j['gpu_launches']))
print(json.dumps(j)[:400])"
timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep device-resident
