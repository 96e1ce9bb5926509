mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_pipelines.py -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/r2ah_pytest.txt; tail -2 gpurun_out/r2ah_pytest.txt
for m in 6 6 4 8; do
  RT_BENCH_COALESCE=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ah_c3_m$m.json 2> gpurun_out/r2ah_c3_m$m.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2ah_c3_m$m.json')); e=j['e2e']
    print('mult $m value',round(j['value']),'e2e',round(e['value']),'tracers',e['tracers_in_flight'],'start us',round(e['host_us_per_start_call']),'wait us',round(e['host_us_waiting_per_call']),j['frame_check']['status'])
except Exception as ex: print('failed', ex); print(open('gpurun_out/r2ah_c3_m$m.err').read()[-600:])
PY
done
