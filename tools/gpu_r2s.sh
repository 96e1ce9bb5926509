mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2s_pytest.txt
tail -4 gpurun_out/r2s_pytest.txt
run() {  # name, config, steps, env...
  name=$1; cfg=$2; steps=$3; shift; shift; shift
  env "$@" timeout 900 python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_$name.json 2> gpurun_out/r2s_bench_$name.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2s_bench_$name.json')); r=j['roofline']
    print('$name', 'value', round(j['value']), 'e2e', round(j['e2e']['value']), 'ms/frame', round(j['ms_per_frame'],4), 'launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), j['frame_check']['status'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2s_bench_$name.err').read()[-1500:])
PY
}
run c4 c4 5 A=1
run c3 c3 20 A=1
run c2 c2 20 A=1
run c1 c1 20 A=1
run c5 c5 2 A=1
