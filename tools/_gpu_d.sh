python -m pytest tests/test_gpu_prune.py tests/test_rdf_s_device.py tests/test_gpu_parity.py tests/test_nranks.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest.log
python tools/probe_nsplit.py > gpurun_out/r2d_nsplit.txt 2>&1
python tools/probe_host_time.py 1 3 > gpurun_out/r2d_host_time.txt 2>&1
tail -4 gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_nsplit.txt; grep "run()" gpurun_out/r2d_host_time.txt
