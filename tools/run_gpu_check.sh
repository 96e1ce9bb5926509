set -x
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_c3_r1a.json
tail -5 gpurun_out/bench_err.log
python bench.py --steps 10 --warmup 3 --config c2 --no-cpu-baseline 2>>gpurun_out/bench_err.log | tee gpurun_out/bench_c2_r1a.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shadow -s 2 -c 2 -o gpurun_out/prof_shadow_r1a python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -o gpurun_out/prof_trace_r1a python bench.py --steps 1 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
