mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 56 -c 2 -o gpurun_out/r2c_async python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
