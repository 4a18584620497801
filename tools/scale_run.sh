#!/bin/bash
# bench.py at N = 2, 4, 8 ranks on one box (gpurun --gpus 8), both arms
tag=${1:-r2z}
o=gpurun_out
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 > $o/${tag}_bench_n$n.json 2> $o/${tag}_bench_n$n.err; echo "n=$n rc=$?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $o/${tag}_bench_reference_n8.json 2> $o/${tag}_bench_reference_n8.err
tail -c 300 $o/${tag}_bench_n8.err
