# A/B of two builds of the library on the default workload: $1 = alternative .so (relative to the repo root)
export FMCMC_BENCH_CFG5=0
for rep in 1 2; do
for lib in "" "$1"; do
  if [ -n "$lib" ]; then export FMCMC_B200_LIB=$PWD/$lib; tag=alt; else unset FMCMC_B200_LIB; tag=base; fi
  python bench.py --no-cpu-baseline > gpurun_out/ab3_$tag.json 2> gpurun_out/ab3_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/ab3_$tag.json')); print('$tag', 'launch %.4f stepping %.4f value %.4g' % (d['roofline']['launch_ms'], d['stepping_only']['ms_per_step'], d['value']))"
done
done
