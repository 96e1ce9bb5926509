mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2i_pytest.txt
tail -4 gpurun_out/r2i_pytest.txt
# node-visit histogram of one c3 launch (stats kernels)
RT_B200_PRINT_HIST=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2i_hist.json 2> gpurun_out/r2i_hist.err
grep -A3 "histogram" gpurun_out/r2i_hist.err | tail -8
# c4 with two frames in flight
for m in 1 2; do
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --pipelines $m --no-cpu-baseline --no-e2e > gpurun_out/r2i_bench_c4_m$m.json 2> gpurun_out/r2i_bench_c4_m$m.err
python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2i_bench_c4_m$m.json')); print('c4 M=$m value',j['value'],'ms/frame',j['ms_per_frame'],j['frame_check']['status'])
except Exception as e: print('c4 failed', e); print(open('gpurun_out/r2i_bench_c4_m$m.err').read()[-800:])
PY
done
