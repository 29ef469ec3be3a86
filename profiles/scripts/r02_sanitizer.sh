# compute-sanitizer on the kernels this session changed: path 4 (prologue over 16 warps, replicated cubic table, four slices per CTA,
# 64-observation blocks), the CTA-per-chain head, programmatic dependent launch
echo "## memcheck: tests/test_gpu_i8.py + tests/test_gpu_session3.py" > gpurun_out/v13_sanitizer.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_i8.py tests/test_gpu_session3.py -m gpu -q -x 2>&1 | tail -6 >> gpurun_out/v13_sanitizer.txt
echo "## racecheck: CTA-per-chain head + slices per CTA + table selection per chain block + wide design matrix (Gaussian K = 128)" >> gpurun_out/v13_sanitizer.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_session3.py tests/test_gpu_i8.py -m gpu -q -k "cta_head_is_bit or slices_per_cta or table_selection or (wide_design and gaussian)" 2>&1 | tail -6 >> gpurun_out/v13_sanitizer.txt
echo "## synccheck: the same selection" >> gpurun_out/v13_sanitizer.txt
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_session3.py tests/test_gpu_i8.py -m gpu -q -k "cta_head_is_bit or slices_per_cta or table_selection or (wide_design and gaussian)" 2>&1 | tail -6 >> gpurun_out/v13_sanitizer.txt
cat gpurun_out/v13_sanitizer.txt
