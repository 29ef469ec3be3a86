# Final single-GPU evidence of round 2 (v13): tests, smoke, every bench line, the reference arm, ncu captures of cfg5 / few.
python -m pytest tests -m gpu -q > gpurun_out/v13_tests.log 2>&1; tail -3 gpurun_out/v13_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v13_smoke.log 2>&1; tail -4 gpurun_out/v13_smoke.log
python bench.py > gpurun_out/v13_bench.json 2> gpurun_out/v13_bench.err
python bench.py --steps 20 --warmup 3 > gpurun_out/v13_bench20.json 2> gpurun_out/v13_bench20.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/v13_ref.json 2> gpurun_out/v13_ref.err
python bench.py --workload few --check-every 0 --no-cpu-baseline > gpurun_out/v13_bench_few.json 2> gpurun_out/v13_bench_few.err
python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/v13_bench_cfg5.json 2> gpurun_out/v13_bench_cfg5.err
for w in cfg4 cfg2 cfg1; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/v13_bench_$w.json 2> gpurun_out/v13_bench_$w.err; done
python - <<'PY'
import json
for f in ("bench","bench20","ref","bench_few","bench_cfg5","bench_cfg4","bench_cfg2","bench_cfg1"):
    try:
        d=json.load(open(f"gpurun_out/v13_{f}.json"))
        r=d.get("roofline") or {}
        print(f, "value %.4g ms/step %.5f launch %s e2e %s frac %s" % (d["value"], d["ms_per_step"], r.get("launch_ms"), (d.get("e2e") or {}).get("value"), r.get("frac")))
    except Exception as e:
        print(f, "ERR", e)
PY
export FMCMC_BENCH_CFG5=0
ncu --set full --import-source on --clock-control none -k regex:tiled_loglik_i8 -s 2 -c 1 -o gpurun_out/v13_cfg5_i8 -f python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline --check-every 0 > gpurun_out/v13_ncu_cfg5.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"tiled_head_adapt_cta|tiled_loglik_mma" -s 20 -c 2 -o gpurun_out/v13_few -f python bench.py --workload few --check-every 0 --steps 20 --warmup 2 --skip-kernel-warmup --no-cpu-baseline > gpurun_out/v13_ncu_few.log 2>&1
ls -la gpurun_out/*.ncu-rep
