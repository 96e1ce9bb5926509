mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/p2p_check.py > gpurun_out/r2g_p2p.txt 2>&1
grep -E "==|Error|error|Traceback|line " gpurun_out/r2g_p2p.txt | head -30
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_c3_n2.json 2> gpurun_out/r2g_bench_c3_n2.err
python - <<'PY'
import json
txt=open('gpurun_out/r2g_bench_c3_n2.json').read()
j=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('value',j['value'],'e2e',j['e2e'],'check',j['frame_check'].get('status'), j['latency'])
PY
