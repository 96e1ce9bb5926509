mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_timed_path.py -x -q -m gpu -k "steal" 2>&1 | tail -8 ) > gpurun_out/r2ab_pytest.txt
tail -5 gpurun_out/r2ab_pytest.txt
run() {  # name, config, steps, env...
  name=$1; cfg=$2; steps=$3; shift; shift; shift
  env "$@" timeout 200 python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2ab_bench_$name.json 2> gpurun_out/r2ab_bench_$name.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2ab_bench_$name.json')); r=j['roofline']
    print('$name', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'launch ms', r['stage_ms_one_launch_alone'], j['frame_check']['status'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2ab_bench_$name.err').read()[-800:])
PY
}
run c3_voted c3 10 RT_B200_TRAV=voted
run c3_steal c3 10 RT_B200_TRAV=steal
run c4_steal c4 3 RT_B200_TRAV=steal
run c2_steal c2 10 RT_B200_TRAV=steal
RT_B200_TRAV=steal timeout 300 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_wave -s 56 -c 7 --csv --log-file gpurun_out/r2ab_ncu_steal.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2ab_ncu.log 2>&1
