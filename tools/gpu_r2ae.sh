N=${1:-2}
mkdir -p gpurun_out
for rep in 1 2 3; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ae_bench_c3_n${N}_$rep.json 2> gpurun_out/r2ae_bench_c3_n${N}_$rep.err
  python - <<PY
import json
try:
    txt=open('gpurun_out/r2ae_bench_c3_n${N}_$rep.json').read()
    j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    e=j['e2e']
    print('n$N rep $rep value',round(j['value']),'e2e',round(e['value']),'start us',round(e['host_us_per_start_call']),'wait us',round(e['host_us_waiting_per_call']),j['frame_check'].get('status'))
except Exception as ex: print('rep $rep failed', ex); print(open('gpurun_out/r2ae_bench_c3_n${N}_$rep.err').read()[-800:])
PY
done
