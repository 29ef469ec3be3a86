#!/bin/bash
# sweep of the path-4 epilogue shapes / slice counts on the cfg3 bench (tuning aid; numbers land in gpurun_out/)
for ns in ${NSLIST:-6}; do for v in ${VLIST:-0 1 2 3 4}; do for dbg in ${DBGLIST:-0}; do
  FMCMC_PATH=4 FMCMC_I8_SLICES=$ns FMCMC_I8_VARIANT=$v FMCMC_I8_DBG=$dbg timeout 200 python bench.py --steps 40 --warmup 3 --skip-kernel-warmup --no-cpu-baseline > gpurun_out/sw_ns${ns}_v${v}_d${dbg}.json 2> gpurun_out/sw_ns${ns}_v${v}_d${dbg}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_ns${ns}_v${v}_d${dbg}.json")); print("NS=$ns variant=$v dbg=$dbg", "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "hot ms %.3f" % d["roofline"]["launch_ms"], "path", d["config"]["path"], "acc %.3f" % d["accept_rate"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("NS=$ns variant=$v dbg=$dbg FAILED", e); print(open("gpurun_out/sw_ns${ns}_v${v}_d${dbg}.err").read()[-600:])
PY
done; done; done
