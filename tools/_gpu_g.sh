python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2g_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2g_ref_n2.json 2> gpurun_out/r2g_ref_n2.err
python -m pytest tests/test_gpu_parity.py tests/test_nranks.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest.log
tail -c 600 gpurun_out/r2g_bench_n2.err; tail -3 gpurun_out/r2g_pytest.log
