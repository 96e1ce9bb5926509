mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2j_pytest.txt
tail -4 gpurun_out/r2j_pytest.txt
for v in split voted; do
  RT_B200_TRAV=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2j_bench_$v.json 2> gpurun_out/r2j_bench_$v.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2j_bench_$v.json')); r=j['roofline']
    print('$v', 'value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), 'alone launch ms', r['stage_ms_one_launch_alone'], 'frac', round(r['frac'],4), j['frame_check']['status'])
except Exception as e: print('$v failed', e); print(open('gpurun_out/r2j_bench_$v.err').read()[-1500:])
PY
done
RT_B200_OCC=6 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2j_bench_split_occ6.json 2> gpurun_out/r2j_bench_split_occ6.err
python -c "
import json; j=json.load(open('gpurun_out/r2j_bench_split_occ6.json')); print('split occ6', round(j['value']), j['roofline']['stage_ms_one_launch_alone'], j['frame_check']['status'])"
timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:k_wave -s 56 -c 7 --csv --log-file gpurun_out/r2j_ncu_split.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1 > gpurun_out/r2j_ncu.log 2>&1
for c in c4 c2; do
timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2j_bench_$c.json 2> gpurun_out/r2j_bench_$c.err
python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2j_bench_$c.json')); print('$c value',j['value'],'ms/frame',j['ms_per_frame'],j['frame_check']['status'])
except Exception as e: print('$c failed', e); print(open('gpurun_out/r2j_bench_$c.err').read()[-800:])
PY
done
