python -m pytest tests -m gpu -x -q > gpurun_out/s8_tests.log 2>&1; tail -3 gpurun_out/s8_tests.log
export FMCMC_BENCH_CFG5=0
python bench.py --no-cpu-baseline > gpurun_out/s8_cfg3.json 2> gpurun_out/s8_cfg3.err
python - <<'PY'
import json
for f in ("cfg3",):
    try:
        d=json.load(open(f"gpurun_out/s8_{f}.json"))
        print(f, "value %.4g ms/step %.4f stepping %.4f launch %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["stepping_only"]["ms_per_step"], d["roofline"]["launch_ms"], d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__inst_executed_op_shared_ld.sum -k regex:tiled_loglik_i8 -c 2 --csv --log-file gpurun_out/s8_ncu_wavefronts.csv python bench.py --steps 2 --warmup 1 --skip-kernel-warmup --no-cpu-baseline > gpurun_out/s8_ncu.log 2>&1
tail -4 gpurun_out/s8_ncu_wavefronts.csv | awk -F'","' '{print $(NF-2), $NF}'
