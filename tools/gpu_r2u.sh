mkdir -p gpurun_out
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_frame -c 1 -o gpurun_out/r2u_kframe_c3 -f python bench.py $B > gpurun_out/r2u_ncu_kframe_c3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_wave" -s 40 -c 10 --csv --page raw --log-file gpurun_out/r2u_kwave_c4_raw.csv python bench.py --config c4 $B > gpurun_out/r2u_ncu_kwave_c4.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_wave" -s 56 -c 7 --csv --page raw --log-file gpurun_out/r2u_kwave_c2_raw.csv python bench.py --config c2 $B > gpurun_out/r2u_ncu_kwave_c2.log 2>&1
ls -la gpurun_out/r2u*
