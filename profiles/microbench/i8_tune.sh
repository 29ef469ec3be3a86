#!/bin/bash
# epilogue experiments on a -DFMCMC_I8_TUNE_HOOKS build (results are garbage, only the time matters).  Build it with
#   (cd fmcmc_b200/csrc && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2 \
#        -DFMCMC_I8_TUNE_HOOKS -shared -cudart static -o ../libfmcmcb200_tune.so fmcmc_b200.cu)
#   1 = no TMEM loads, 2 = no epilogue arithmetic, 4 = no softplus-table LDS, 8 = no I2F (XU) in the reassembly
for t in ${TLIST:-0 4 8 12 5 13}; do
  FMCMC_B200_LIB=$PWD/fmcmc_b200/libfmcmcb200_tune.so FMCMC_I8_TUNE=$t timeout 200 python bench.py --steps 40 --warmup 3 --skip-kernel-warmup --no-cpu-baseline ${WL:-} > gpurun_out/tune_$t.json 2> gpurun_out/tune_$t.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/tune_$t.json')); print('tune=$t hot ms %.3f' % d['roofline']['launch_ms'])
except Exception as e: print('tune=$t FAILED', e, open('gpurun_out/tune_$t.err').read()[-300:])"
done
