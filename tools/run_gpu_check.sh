timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --config c2 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_c2_n1.json | cut -c1-300
python bench.py --steps 20 --warmup 3 --config c1 2>/dev/null | tee gpurun_out/bench_c1_n1.json | cut -c1-300
python bench.py --steps 10 --warmup 3 --config c4 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_c4_n1.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/bench_ref_c3.json | cut -c1-400
