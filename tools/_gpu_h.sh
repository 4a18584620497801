python tools/bench_kernel.py 4 > gpurun_out/r2h_plain4.txt 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pair_prune -s 4 -c 1 -f -o gpurun_out/r2h_tri python tools/bench_kernel.py 4 > gpurun_out/r2h_ncu_tri.log 2>&1
python tools/bench_kernel.py 2 > gpurun_out/r2h_plain2.txt 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pair_prune -s 4 -c 1 -f -o gpurun_out/r2h_ortho python tools/bench_kernel.py 2 > gpurun_out/r2h_ncu_ortho.log 2>&1
cat gpurun_out/r2h_plain4.txt gpurun_out/r2h_plain2.txt
