# round 2, GPU call A: full GPU test suite, bench (both arms), launch list and a full-set capture of HEAD's k_wave batch
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
nproc >> gpurun_out/r2a_gpu.txt
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2a_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c3_n1.json 2> gpurun_out/r2a_bench_c3_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref_c3.json 2> gpurun_out/r2a_bench_ref_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2a_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 28 -c 7 -o gpurun_out/r2a_kwave python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_pytest.txt; head -c 1500 gpurun_out/r2a_bench_c3_n1.json; tail -3 gpurun_out/r2a_bench_c3_n1.err; head -c 800 gpurun_out/r2a_bench_ref_c3.json
