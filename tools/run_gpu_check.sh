python bench.py --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_n1_r1g.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log; cut -c1-1500 gpurun_out/bench_c3_n1_r1g.json
python bench.py --steps 60 --warmup 3 --no-cpu-baseline --pipelines 1 2>&1 | cut -c1-400
