mkdir -p gpurun_out
RT_B200_COALESCE_TRACE=1 RT_BENCH_COALESCE=6 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2ag_c3_m6.json 2> gpurun_out/r2ag_c3_m6.err
grep -c coalesce gpurun_out/r2ag_c3_m6.err
grep coalesce gpurun_out/r2ag_c3_m6.err | tail -40 | cut -c1-220
python - <<PY
import json
j=json.load(open('gpurun_out/r2ag_c3_m6.json')); e=j['e2e']
print('e2e',round(e['value']),'start us',round(e['host_us_per_start_call']),'wait us',round(e['host_us_waiting_per_call']))
PY
