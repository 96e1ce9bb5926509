mkdir -p gpurun_out
run() {  # name, config, steps, env...
  name=$1; cfg=$2; steps=$3; shift; shift; shift
  env "$@" timeout 900 python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench_$name.json 2> gpurun_out/r2q_bench_$name.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2q_bench_$name.json')); r=j['roofline']
    print('$name', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), j['frame_check']['status'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2q_bench_$name.err').read()[-1500:])
PY
}
run c4_off c4 3 RT_B200_BIN=0
for k in 1 8 2 3 7 11; do run c4_b3_key$k c4 3 RT_B200_BIN=3 RT_B200_BIN_KEY=$k; done
for k in 4 7; do run c4_b5_key$k c4 3 RT_B200_BIN=5 RT_B200_BIN_KEY=$k; done
run c3_b3_key3 c3 10 RT_B200_BIN=3 RT_B200_BIN_KEY=3
run c3_b5_key7 c3 10 RT_B200_BIN=5 RT_B200_BIN_KEY=7
