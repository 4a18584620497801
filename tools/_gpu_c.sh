python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest.log
python tools/probe_host_time.py 1 3 > gpurun_out/r2c_host_time.txt 2>&1
python bench.py --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" >> gpurun_out/r2c_bench.err
tail -4 gpurun_out/r2c_pytest.log; grep "run()" gpurun_out/r2c_host_time.txt; tail -c 300 gpurun_out/r2c_bench.err
