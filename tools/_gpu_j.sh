for v in default dw16 dw4 t192 t192dw16; do
  if [ $v = default ]; then unset RDFK_LIB; else export RDFK_LIB=$PWD/pmda_b200/csrc/variants/$v.so; fi
  python tools/bench_kernel.py 4 >> gpurun_out/r2j_variants.txt 2>&1
done
for v in default o_iat o_192 o_192iat; do
  if [ $v = default ]; then unset RDFK_LIB; else export RDFK_LIB=$PWD/pmda_b200/csrc/variants/$v.so; fi
  python tools/bench_kernel.py 2 1 >> gpurun_out/r2j_variants.txt 2>&1
done
cat gpurun_out/r2j_variants.txt
