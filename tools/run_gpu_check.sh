timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_c3_n1.json | cut -c1-150
for n in 2 4; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 2>/dev/null | tee gpurun_out/bench_c3_n$n.json | cut -c1-150; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 --config c4 2>/dev/null | tee gpurun_out/bench_c4_n4.json | cut -c1-150
