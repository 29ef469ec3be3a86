python -m pytest tests -m gpu -x -q > gpurun_out/s10_tests.log 2>&1; tail -3 gpurun_out/s10_tests.log
python bench.py --no-cpu-baseline > gpurun_out/s10_cfg3.json 2> gpurun_out/s10_cfg3.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s10_cfg3.json"))
print("cfg3 value %.4g ms/step %.4f stepping %.4f launch %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["stepping_only"]["ms_per_step"], d["roofline"]["launch_ms"], d["e2e"]["value"]))
print("cfg5", json.dumps(d.get("cfg5"))[:1500])
PY
