export FMCMC_B200_LIB=$PWD/fmcmc_b200/libfmcmcb200_b48.so
timeout 600 python -m pytest tests/test_gpu_i8.py tests/test_gpu_fullsize.py tests/test_gpu_session3.py -m gpu -x -q 2>&1 | tail -3
unset FMCMC_B200_LIB
for rep in 1 2; do
for lib in "" fmcmc_b200/libfmcmcb200_b48.so; do
  if [ -n "$lib" ]; then export FMCMC_B200_LIB=$PWD/$lib; tag=b48; else unset FMCMC_B200_LIB; tag=base; fi
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b48_$tag.json 2> gpurun_out/b48_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/b48_$tag.json')); c=d.get('cfg5') or {}; print('$tag', 'launch %.4f stepping %.4f value %.4g | cfg5 launch %s ms/step %s' % (d['roofline']['launch_ms'], d['stepping_only']['ms_per_step'], d['value'], (c.get('roofline') or {}).get('launch_ms'), c.get('ms_per_step')), d['timed_region']['hot_launch_ms_series'][:12])"
done
done
