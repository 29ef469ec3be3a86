python -m pytest tests -m gpu -x -q > gpurun_out/s12_tests.log 2>&1; tail -3 gpurun_out/s12_tests.log
export FMCMC_BENCH_CFG5=0
for hc in 0 1; do
  FMCMC_HEAD_CTA=$hc python bench.py --no-cpu-baseline --workload few --check-every 0 > gpurun_out/s12_few_hc$hc.json 2> gpurun_out/s12_few_hc$hc.err
  python -c "
import json
d=json.load(open('gpurun_out/s12_few_hc$hc.json')); print('few head_cta=$hc', 'value %.4g ms/step %.4f stepping %.4f launch %.4f e2e %.4g accept %.4f' % (d['value'], d['ms_per_step'], d['stepping_only']['ms_per_step'], d['roofline']['launch_ms'], d['e2e']['value'], d['accept_rate']), d['timed_region']['repeat_ms_per_step'])"
done
