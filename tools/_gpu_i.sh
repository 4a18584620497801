python -m pytest tests/test_gpu_prune.py tests/test_gpu_fullsize.py tests/test_gpu_check.py tests/test_golden_counts.py tests/test_pinning.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest.log
python tools/bench_kernel.py 4 2 > gpurun_out/r2i_kernel.txt 2>&1
tail -3 gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_kernel.txt
