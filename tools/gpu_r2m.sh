mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_timed_path.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2m_pytest.txt
tail -3 gpurun_out/r2m_pytest.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_$name.json 2> gpurun_out/r2m_bench_$name.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2m_bench_$name.json')); r=j['roofline']
    print('$name', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), j['frame_check']['status'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2m_bench_$name.err').read()[-1500:])
PY
}
for t in 1 2 4 8 64; do run defer$t RT_B200_TRAV=defer RT_B200_DQ=$t; done
for t in 1 4; do
RT_B200_DQ=$t RT_B200_TRAV=defer timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_wave -s 56 -c 7 --csv --log-file gpurun_out/r2m_ncu_defer$t.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2m_ncu_defer$t.log 2>&1
done
