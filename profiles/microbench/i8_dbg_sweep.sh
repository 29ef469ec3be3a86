#!/bin/bash
for v in 0 2; do for dbg in 0 1 2 3; do
  FMCMC_PATH=4 FMCMC_I8_VARIANT=$v FMCMC_I8_DBG=$dbg timeout 200 python bench.py --steps 30 --warmup 3 --skip-kernel-warmup --no-cpu-baseline > gpurun_out/dbg_v${v}_d${dbg}.json 2> gpurun_out/dbg_v${v}_d${dbg}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/dbg_v${v}_d${dbg}.json")); print("variant=$v dbg=$dbg", "hot ms %.3f" % d["roofline"]["launch_ms"])
except Exception as e:
    print("variant=$v dbg=$dbg FAILED", e); print(open("gpurun_out/dbg_v${v}_d${dbg}.err").read()[-400:])
PY
done; done
