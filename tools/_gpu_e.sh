python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest.log
python tools/probe_host_time.py 1 3 > gpurun_out/r2e_host_time.txt 2>&1
tail -3 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_host_time.txt
