mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2k_pytest.txt
tail -6 gpurun_out/r2k_pytest.txt
for v in bvh8 bvh4; do
  if [ $v = bvh4 ]; then export RT_B200_LIBDIR=$PWD/raytrace_b200/lib_bvh4; else unset RT_B200_LIBDIR; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_$v.json 2> gpurun_out/r2k_bench_$v.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2k_bench_$v.json')); r=j['roofline']
    print('$v', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'alone', j['latency']['ms_per_frame_alone'], 'launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), 'nodes/ray', r['nodes_per_ray'], r['tri_tests_per_ray'], j['frame_check']['status'], j['build'])
except Exception as e: print('$v failed', e); print(open('gpurun_out/r2k_bench_$v.err').read()[-1500:])
PY
  timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_wave -s 56 -c 7 --csv --log-file gpurun_out/r2k_ncu_$v.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2k_ncu_$v.log 2>&1
done
unset RT_B200_LIBDIR
for c in c4; do
timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_bench_$c.json 2> gpurun_out/r2k_bench_$c.err
python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2k_bench_$c.json')); print('$c value',j['value'],'ms/frame',j['ms_per_frame'],j['frame_check']['status'])
except Exception as e: print('$c failed', e); print(open('gpurun_out/r2k_bench_$c.err').read()[-800:])
PY
done
