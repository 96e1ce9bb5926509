mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_timed_path.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2p_pytest.txt
tail -3 gpurun_out/r2p_pytest.txt
run() {  # name, config, steps, env...
  name=$1; cfg=$2; steps=$3; shift; shift; shift
  env "$@" timeout 900 python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_$name.json 2> gpurun_out/r2p_bench_$name.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2p_bench_$name.json')); r=j['roofline']
    print('$name', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), j['frame_check']['status'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2p_bench_$name.err').read()[-1500:])
PY
}
run c4_off c4 3 RT_B200_BIN=0
for b in 3 4 5 6; do run c4_bin$b c4 3 RT_B200_BIN=$b; done
run c3_off c3 10 RT_B200_BIN=0
for b in 4 5; do run c3_bin$b c3 10 RT_B200_BIN=$b; done
run c2_bin4 c2 10 RT_B200_BIN=4
RT_B200_BIN=5 timeout 900 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_wave|k_bin" -s 96 -c 40 --csv --log-file gpurun_out/r2p_ncu_c4.csv python bench.py --config c4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2p_ncu_c4.log 2>&1
