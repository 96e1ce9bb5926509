N=${1:-4}
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config c3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/r2z_bench_${name}_n$N.json 2> gpurun_out/r2z_bench_${name}_n$N.err
  python - <<PY
import json
try:
    txt=open('gpurun_out/r2z_bench_${name}_n$N.json').read()
    j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print('$name n$N value',round(j['value']),'check',j['frame_check'].get('status'), [(round(r['ms_per_step'],3), r['rays_per_step']) for r in j['per_rank']])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r2z_bench_${name}_n$N.err').read()[-1500:])
PY
}
run bp1 A=1
run bp0 RT_BENCH_BACKPRESSURE=0
EXTRA="--shard-order modulo" run modulo A=1
EXTRA="--gather nccl" run nccl A=1
