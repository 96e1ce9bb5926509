# final validation at HEAD: what the driver runs (GPU tests, smoke, bench both arms) + the evidence captures bench.py quotes
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/final_pytest.txt; tail -3 gpurun_out/final_pytest.txt
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/final_smoke.txt; cat gpurun_out/final_smoke.txt
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wave -s 56 -c 7 -o gpurun_out/final_kwave_c3 -f python bench.py $B > gpurun_out/final_ncu_kwave_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_wave" -s 40 -c 10 --csv --page raw --log-file gpurun_out/final_kwave_c4_raw.csv python bench.py --config c4 $B > gpurun_out/final_ncu_kwave_c4.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_wave" -s 56 -c 7 --csv --page raw --log-file gpurun_out/final_kwave_c2_raw.csv python bench.py --config c2 $B > gpurun_out/final_ncu_kwave_c2.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref_c3.json 2> gpurun_out/final_bench_ref_c3.err; cut -c1-300 gpurun_out/final_bench_ref_c3.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_c3_n1.json 2> gpurun_out/final_bench_c3_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/final_bench_c3_n1.json'))
print('c3 value',round(j['value']),'e2e',round(j['e2e']['value']),j['frame_check']['status'],'frac',round(j['roofline']['frac'],4),'traffic',j['roofline']['traffic'],j['roofline'].get('traffic_source'),'cpu',j['cpu_baseline']['value'],'launches',j['gpu_launches'],j['clocks'])
PY
ls -la gpurun_out/final*
