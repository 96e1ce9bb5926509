mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2h_pytest.txt
tail -6 gpurun_out/r2h_pytest.txt
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/r2h_bench_c4_n1.json 2> gpurun_out/r2h_bench_c4_n1.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2h_bench_c4_n1.json'))
    print('c4 value',j['value'],'ms/frame',j['ms_per_frame'],'e2e',j['e2e']['value'],'check',j['frame_check'],'frac',j['roofline']['frac'], j['roofline']['stage_ms_one_launch_alone'], j.get('cpu_baseline'))
except Exception as e: print('c4 failed', e); print(open('gpurun_out/r2h_bench_c4_n1.err').read()[-1500:])
PY
