# round 2, GPU call E: supersampling + tile window tests, c5 and c3 bench lines
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2e_pytest.txt
tail -5 gpurun_out/r2e_pytest.txt
timeout 900 python bench.py --config c5 --steps 2 --warmup 1 > gpurun_out/r2e_bench_c5_n1.json 2> gpurun_out/r2e_bench_c5_n1.err
head -c 1800 gpurun_out/r2e_bench_c5_n1.json; tail -3 gpurun_out/r2e_bench_c5_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench_c3_n1.json 2> gpurun_out/r2e_bench_c3_n1.err
head -c 600 gpurun_out/r2e_bench_c3_n1.json; tail -3 gpurun_out/r2e_bench_c3_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2e_bench_ref_c3.json 2> gpurun_out/r2e_bench_ref_c3.err
head -c 300 gpurun_out/r2e_bench_ref_c3.json
