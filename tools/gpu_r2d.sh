# round 2, GPU call D: FMA slab test + packed stack slots -- parity, A/B
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2d_pytest.txt
tail -3 gpurun_out/r2d_pytest.txt
for v in lib lib_nofma; do
  RT_B200_LIBDIR=raytrace_b200/$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench_$v.json 2> gpurun_out/r2d_bench_$v.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2d_bench_$v.json')); r=j['roofline']
    print('$v', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'alone launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), j['frame_check']['status'], 'alone frame', j['latency']['ms_per_frame_alone'])
except Exception as e: print('$v failed', e); print(open('gpurun_out/r2d_bench_$v.err').read()[-1500:])
PY
done
timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:k_wave -s 56 -c 7 --csv --log-file gpurun_out/r2d_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2d_ncu.log 2>&1
