mkdir -p gpurun_out
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1"
timeout 600 ncu --set full --clock-control none -k regex:k_wave -s 56 -c 7 --csv --page raw --log-file gpurun_out/r2ad_kwave_c3_b8_raw.csv python bench.py $B > gpurun_out/r2ad_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ad_bench_c3_n1.json 2> gpurun_out/r2ad_bench_c3_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2ad_bench_c3_n1.json'))
print('c3 value',round(j['value']),'e2e',round(j['e2e']['value']),j['e2e']['tracers_in_flight'],j['frame_check']['status'],'frac',round(j['roofline']['frac'],4),j['run']['frames_per_launch'],j['run']['launches_in_flight'],j['roofline']['stage_ms_one_launch_alone'])
PY
