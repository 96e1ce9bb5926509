mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 900 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2o_bench_$name.json 2> gpurun_out/r2o_bench_$name.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2o_bench_$name.json')); r=j['roofline']
    print('$name', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), j['frame_check']['status'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2o_bench_$name.err').read()[-1500:])
PY
}
run voted RT_B200_TRAV=voted
for t in 2 4 8 64; do run defer$t RT_B200_TRAV=defer RT_B200_DQ=$t; done
run split RT_B200_TRAV=split
