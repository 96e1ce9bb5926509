timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_c3_n1.json | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e'], j['roofline']['stage_ms'])"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --config c2 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e'], j['roofline']['stage_ms'])"
