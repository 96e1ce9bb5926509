set -x
python bench.py --steps 100 --warmup 3 > gpurun_out/r1h_bench_c3_n1.json 2> gpurun_out/r1h_bench_c3_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1h_bench_ref_c3.json 2>/dev/null
python bench.py --config c2 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_bench_c2_n1.json 2>/dev/null
python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_bench_c4_n1.json 2>/dev/null
python bench.py --config c1 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_bench_c1_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r1h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame -s 4 -c 1 -o gpurun_out/prof_frame_r1h -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 > gpurun_out/ncu_full_r1h.log 2>&1
for f in gpurun_out/r1h_bench_*.json; do echo $f; cut -c1-260 $f; done
