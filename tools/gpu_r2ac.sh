mkdir -p gpurun_out
for bm in "4 2" "4 3" "8 2" "2 4" "6 2" "8 3" "16 1"; do
  set -- $bm
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --batch $1 --pipelines $2 > gpurun_out/r2ac_b$1_m$2.json 2> gpurun_out/r2ac_b$1_m$2.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2ac_b$1_m$2.json')); print('B $1 M $2 value', round(j['value']), 'ms/frame', round(j['ms_per_frame'],4), j['run']['frames_per_launch'], j['run']['launches_in_flight'], j['frame_check']['status'])
except Exception as e: print('B $1 M $2 failed', e); print(open('gpurun_out/r2ac_b$1_m$2.err').read()[-500:])
PY
done
