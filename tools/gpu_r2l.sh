mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2l_pytest.txt
tail -6 gpurun_out/r2l_pytest.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2l_bench_$name.json 2> gpurun_out/r2l_bench_$name.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2l_bench_$name.json')); r=j['roofline']
    print('$name', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), j['frame_check']['status'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2l_bench_$name.err').read()[-1500:])
PY
}
run voted RT_B200_TRAV=voted
run defer32 RT_B200_TRAV=defer RT_B200_DQ=32
run defer16 RT_B200_TRAV=defer RT_B200_DQ=16
run defer24 RT_B200_TRAV=defer RT_B200_DQ=24
run defer48 RT_B200_TRAV=defer RT_B200_DQ=48
run defer8 RT_B200_TRAV=defer RT_B200_DQ=8
RT_B200_TRAV=defer timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_wave -s 56 -c 7 --csv --log-file gpurun_out/r2l_ncu_defer.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2l_ncu_defer.log 2>&1
