# evidence capture at HEAD: launch list + ncu --set full of the traversal kernels (c3 batch), the whole-frame kernels, shade/resolve
mkdir -p gpurun_out
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2t_launches_c3.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_wave -s 56 -c 7 -o gpurun_out/r2t_kwave_c3 -f python bench.py $B > gpurun_out/r2t_ncu_kwave_c3.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_frame|k_shade|k_resolve" -s 3 -c 3 -o gpurun_out/r2t_kframe_c3 -f python bench.py $B > gpurun_out/r2t_ncu_kframe_c3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_shade|k_resolve" -s 16 -c 2 -o gpurun_out/r2t_post_c3 -f python bench.py $B > gpurun_out/r2t_ncu_post_c3.log 2>&1
ls -la gpurun_out/r2t*
