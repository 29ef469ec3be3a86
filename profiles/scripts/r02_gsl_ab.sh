timeout 300 python -m pytest tests/test_gpu_i8.py tests/test_gpu_fullsize.py tests/test_gpu_session3.py -m gpu -x -q > gpurun_out/s13_tests.log 2>&1; tail -3 gpurun_out/s13_tests.log
export FMCMC_BENCH_CFG5=0
for rep in 1 2; do
for g in 1 2 4; do
  FMCMC_I8_GSL=$g timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s13_gsl$g.json 2> gpurun_out/s13_gsl$g.err
  python -c "
import json; d=json.load(open('gpurun_out/s13_gsl$g.json')); print('gsl=$g', 'launch %.4f stepping %.4f value %.4g e2e %.4g' % (d['roofline']['launch_ms'], d['stepping_only']['ms_per_step'], d['value'], d['e2e']['value']))"
done
done
