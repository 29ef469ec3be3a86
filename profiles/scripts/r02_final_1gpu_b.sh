# Final single-GPU lines of round 2 (v14 = v13 kernels + lazy kernel views / mcmc.list, L2 hints in the few-chain regime)
python -m pytest tests -m gpu -q > gpurun_out/v14_tests.log 2>&1; tail -3 gpurun_out/v14_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v14_smoke.log 2>&1; tail -4 gpurun_out/v14_smoke.log
python bench.py > gpurun_out/v14_bench.json 2> gpurun_out/v14_bench.err
python bench.py --steps 20 --warmup 3 > gpurun_out/v14_bench20.json 2> gpurun_out/v14_bench20.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/v14_ref.json 2> gpurun_out/v14_ref.err
python bench.py --workload few --check-every 0 --no-cpu-baseline > gpurun_out/v14_bench_few.json 2> gpurun_out/v14_bench_few.err
python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/v14_bench_cfg5.json 2> gpurun_out/v14_bench_cfg5.err
for w in cfg4 cfg2 cfg1; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/v14_bench_$w.json 2> gpurun_out/v14_bench_$w.err; done
python - <<'PY'
import json
for f in ("bench","bench20","ref","bench_few","bench_cfg5","bench_cfg4","bench_cfg2","bench_cfg1"):
    try:
        d=json.load(open(f"gpurun_out/v14_{f}.json"))
        r=d.get("roofline") or {}
        print(f, "value %.4g ms/step %.5f stepping %s launch %s e2e %s frac %s" % (d["value"], d["ms_per_step"], (d.get("stepping_only") or {}).get("ms_per_step"), r.get("launch_ms"), (d.get("e2e") or {}).get("value"), r.get("frac")))
    except Exception as e:
        print(f, "ERR", e)
PY
