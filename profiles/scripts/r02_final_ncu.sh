export FMCMC_BENCH_CFG5=0
ncu --set full --import-source on --clock-control none -k regex:tiled_loglik_i8 -s 3 -c 1 -o gpurun_out/v13_cfg3_i8 -f python bench.py --steps 3 --warmup 1 --skip-kernel-warmup --no-cpu-baseline > gpurun_out/v13_ncu.log 2>&1
ls -la gpurun_out/v13_cfg3_i8.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v13_launches.csv python bench.py --steps 20 --warmup 3 --skip-kernel-warmup --no-cpu-baseline > gpurun_out/v13_launches.log 2>&1
tail -3 gpurun_out/v13_launches.csv | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/v13_few_launches.csv python bench.py --workload few --check-every 0 --steps 40 --warmup 3 --skip-kernel-warmup --no-cpu-baseline > gpurun_out/v13_few_launches.log 2>&1
tail -3 gpurun_out/v13_few_launches.csv | cut -c1-300
