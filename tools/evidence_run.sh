#!/bin/bash
# The evidence run of a round (one gpurun call, 1 GPU): everything
# tools/collect_evidence.py turns into tracked files under profiles/.
#   gpurun --timeout 1500 -- 'bash tools/evidence_run.sh r2z'
tag=${1:-r2z}
o=gpurun_out
python -m pytest tests -m gpu -q > $o/${tag}_pytest.log 2>&1; echo "rc=$?" >> $o/${tag}_pytest.log
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "rc=$?" >> $o/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err
python tools/cpu_configs.py > $o/${tag}_cpu_configs.txt 2>&1
python tools/bench_contacts.py > $o/${tag}_contacts.txt 2>&1
python tools/bench_hbonds.py > $o/${tag}_hbonds.txt 2>&1
python tools/probe_host_time.py 1 3 > $o/${tag}_host_time.txt 2>&1
python tools/bench_full_length.py 2 3 4 5 > $o/${tag}_full_length.txt 2>&1
# ncu: each only after the same command exited 0 without it
python tools/bench_kernel.py 4 > $o/${tag}_plain4.txt 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pair_prune -s 4 -c 1 -f -o $o/${tag}_tri python tools/bench_kernel.py 4 > $o/${tag}_ncu_tri.log 2>&1
python tools/bench_kernel.py 2 > $o/${tag}_plain2.txt 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pair_prune -s 4 -c 1 -f -o $o/${tag}_ortho python tools/bench_kernel.py 2 > $o/${tag}_ncu_ortho.log 2>&1
echo 64 > $o/${tag}_ortho_frames.txt
python tools/probe_sites.py 400 > $o/${tag}_plain_sites.txt 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sites_kernel -s 8 -c 1 -f -o $o/${tag}_sites python tools/probe_sites.py 400 > $o/${tag}_ncu_sites.log 2>&1
python bench.py --steps 2 --warmup 3 --frames-per-step 32 --no-cpu-baseline --no-configs > $o/${tag}_bench_small.json 2> $o/${tag}_bench_small.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_bench_small.csv python bench.py --steps 2 --warmup 3 --frames-per-step 32 --no-cpu-baseline --no-configs > $o/${tag}_ncu_launch.log 2>&1
tail -2 $o/${tag}_pytest.log; tail -1 $o/${tag}_bench.err; cat $o/${tag}_plain4.txt $o/${tag}_plain2.txt
