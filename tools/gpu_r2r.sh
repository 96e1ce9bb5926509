mkdir -p gpurun_out
run() {  # name, config, steps, env...
  name=$1; cfg=$2; steps=$3; shift; shift; shift
  env "$@" timeout 900 python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2r_bench_$name.json 2> gpurun_out/r2r_bench_$name.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2r_bench_$name.json')); r=j['roofline']
    print('$name', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), j['frame_check']['status'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2r_bench_$name.err').read()[-1500:])
PY
}
for b in 1 2 4; do run c4_b${b}_key7 c4 3 RT_B200_BIN=$b RT_B200_BIN_KEY=7; done
run c4_b3_key5 c4 3 RT_B200_BIN=3 RT_B200_BIN_KEY=5
run c4_b2_key5 c4 3 RT_B200_BIN=2 RT_B200_BIN_KEY=5
run c4_b3_key7_defer c4 3 RT_B200_BIN=3 RT_B200_BIN_KEY=7 RT_B200_TRAV=defer RT_B200_DQ=8
run c2_off c2 10 RT_B200_BIN=0
run c2_b3 c2 10 RT_B200_BIN=3
run c2_b2 c2 10 RT_B200_BIN=2
run c3_b1 c3 10 RT_B200_BIN=1
run c3_b2 c3 10 RT_B200_BIN=2
