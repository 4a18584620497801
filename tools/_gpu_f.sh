python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest.log
python tools/probe_host_time.py 1 3 > gpurun_out/r2f_host_time.txt 2>&1
python bench.py --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?" >> gpurun_out/r2f_bench.err
tail -3 gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_host_time.txt; tail -c 300 gpurun_out/r2f_bench.err
