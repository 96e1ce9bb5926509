mkdir -p gpurun_out
run() {  # name, config, steps, env...
  name=$1; cfg=$2; steps=$3; shift; shift; shift
  env "$@" timeout 300 python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2x_bench_$name.json 2> gpurun_out/r2x_bench_$name.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2x_bench_$name.json')); r=j['roofline']
    print('$name', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'launch ms', r['stage_ms_one_launch_alone'], 'nodes/ray', round(r['nodes_per_ray'],2), round(r['tri_tests_per_ray'],2), j['frame_check']['status'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2x_bench_$name.err').read()[-800:])
PY
}
run c3_voted_l2 c3 10 RT_B200_TRAV=voted
run c3_voted_l4 c3 10 RT_B200_TRAV=voted RT_B200_LEAF_SIZE=4
run c3_defer_l4 c3 10 RT_B200_TRAV=defer RT_B200_DQ=8 RT_B200_LEAF_SIZE=4
run c3_defer_l8 c3 10 RT_B200_TRAV=defer RT_B200_DQ=8 RT_B200_LEAF_SIZE=8
run c3_defer_l4_q64 c3 10 RT_B200_TRAV=defer RT_B200_DQ=64 RT_B200_LEAF_SIZE=4
run c4_defer_l4 c4 3 RT_B200_TRAV=defer RT_B200_DQ=8 RT_B200_LEAF_SIZE=4
run c4_voted_l4 c4 3 RT_B200_TRAV=voted RT_B200_LEAF_SIZE=4
