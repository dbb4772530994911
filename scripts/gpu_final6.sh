mkdir -p gpurun_out/final
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 ) | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py ) > gpurun_out/final/bench_cfg2.json 2> gpurun_out/final/bench_cfg2.err
grep real gpurun_out/final/bench_cfg2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref_cfg2.json 2> gpurun_out/final/bench_ref_cfg2.err
for w in cfg1 cfg3 cfg4-adaptive; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/final/bench_*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f,"ERR",e); continue
    cb=j.get("cpu_baseline") or {}; e2e=j.get("e2e") or {}
    print(f.split("/")[-1], "value %.4g e2e %.4g cpu %s" % (j["value"], e2e.get("value",0), cb.get("value")))
PY
