python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_c3_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/bench_err2.log | tee gpurun_out/bench_c3_n2.json
