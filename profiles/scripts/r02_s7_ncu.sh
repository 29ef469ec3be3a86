export FMCMC_BENCH_CFG5=0
ncu --set full --import-source on --clock-control none -k regex:tiled_loglik_i8 -s 3 -c 1 -o gpurun_out/s9_cfg3_i8 -f python bench.py --steps 3 --warmup 1 --skip-kernel-warmup --no-cpu-baseline > gpurun_out/s9_ncu.log 2>&1
ls -la gpurun_out/s9_cfg3_i8.ncu-rep
