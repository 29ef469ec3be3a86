for t in 0 1 2 3; do
  FMCMC_B200_LIB=$PWD/fmcmc_b200/libfmcmcb200_tune.so FMCMC_I8_TUNE=$t timeout 300 python bench.py --workload cfg5 --steps 3 --warmup 1 --skip-kernel-warmup --no-cpu-baseline --check-every 0 > gpurun_out/tune5_$t.json 2> gpurun_out/tune5_$t.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/tune5_$t.json')); print('cfg5 tune=$t hot ms %.2f' % d['roofline']['launch_ms'])
except Exception as e: print('tune=$t FAILED', e, open('gpurun_out/tune5_$t.err').read()[-300:])"
done
