timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
RT_B200_SCHED=frame timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
RT_B200_SCHED=waves timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 3 2>/dev/null | tee gpurun_out/bench_c3_n1.json | cut -c1-200
python bench.py --steps 20 --warmup 3 --config c2 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_c2_n1.json | cut -c1-200
python bench.py --steps 20 --warmup 3 --config c1 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_c1_n1.json | cut -c1-200
python bench.py --steps 10 --warmup 3 --config c4 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_c4_n1.json | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | tee gpurun_out/bench_c3_n2.json | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --config c4 2>/dev/null | tee gpurun_out/bench_c4_n2.json | cut -c1-200
