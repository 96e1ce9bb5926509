# round 2, GPU call F (2 GPUs): multi-GPU parity test, c3 bench at N=2 with coalesced start() calls traced
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_timed_path.py -x -q -m gpu -k "multi or p2p or supersampling or tile_window" 2>&1 | tail -15 ) > gpurun_out/r2f_pytest.txt
tail -4 gpurun_out/r2f_pytest.txt
RT_B200_COALESCE_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_c3_n2.json 2> gpurun_out/r2f_bench_c3_n2.err
head -c 400 gpurun_out/r2f_bench_c3_n2.json; echo; grep -c coalesce gpurun_out/r2f_bench_c3_n2.err; grep coalesce gpurun_out/r2f_bench_c3_n2.err | tail -12
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2f_bench_c3_n2.json'))
print('value',j['value'],'e2e',j['e2e'],'check',j['frame_check'].get('status'), j['latency'])
PY
