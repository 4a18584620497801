for v in default v0 v1b v2 v3; do
  if [ $v = default ]; then unset RDFK_LIB; else export RDFK_LIB=$PWD/pmda_b200/csrc/variants/$v.so; fi
  python tools/bench_kernel.py 4 2 >> gpurun_out/r2b_variants.txt 2>&1
done
unset RDFK_LIB
python -m pytest tests/test_gpu_prune.py tests/test_gpu_check.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest.log
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" >> gpurun_out/r2b_bench.err
cat gpurun_out/r2b_variants.txt; tail -3 gpurun_out/r2b_pytest.log; tail -c 400 gpurun_out/r2b_bench.err
