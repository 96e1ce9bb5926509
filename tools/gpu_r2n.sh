mkdir -p gpurun_out
timeout 900 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_wave|k_combine|k_shade" -s 48 -c 24 --csv --log-file gpurun_out/r2n_ncu_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2n_ncu_c4.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_wave -s 42 -c 2 -o gpurun_out/r2n_c4_wave23 -f python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2n_ncu_c4_full.log 2>&1
ls -la gpurun_out/r2n*
