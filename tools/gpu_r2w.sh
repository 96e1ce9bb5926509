# N ranks: the driver's bench command for c3 (and c5), no p2p_check
N=${1:-8}
mkdir -p gpurun_out
for cfg in ${CFGS:-c3 c5}; do
  steps=20; [ $cfg = c5 ] && steps=2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $cfg --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/r2w_bench_${cfg}_n$N.json 2> gpurun_out/r2w_bench_${cfg}_n$N.err
  python - <<PY
import json
try:
    txt=open('gpurun_out/r2w_bench_${cfg}_n$N.json').read()
    j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print('$cfg n$N value',round(j['value']),'e2e',round(j['e2e']['value']),'check',j['frame_check'].get('status'), j.get('latency',{}).get('ms_per_frame_alone'), 'slowest', j['run'].get('slowest_rank') if 'run' in j else None, [round(r['ms_per_step'],3) for r in j['per_rank']])
except Exception as e: print('$cfg failed', e); print(open('gpurun_out/r2w_bench_${cfg}_n$N.err').read()[-1500:])
PY
done
