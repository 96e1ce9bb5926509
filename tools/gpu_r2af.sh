mkdir -p gpurun_out
for m in 8 6; do
  RT_BENCH_COALESCE=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2af_c3_m$m.json 2> gpurun_out/r2af_c3_m$m.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2af_c3_m$m.json')); e=j['e2e']
    print('mult $m value',round(j['value']),'e2e',round(e['value']),'tracers',e['tracers_in_flight'],'start us',round(e['host_us_per_start_call']),'wait us',round(e['host_us_waiting_per_call']))
except Exception as ex: print('failed', ex); print(open('gpurun_out/r2af_c3_m$m.err').read()[-600:])
PY
done
