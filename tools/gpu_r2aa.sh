mkdir -p gpurun_out
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pipelines 1"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wave -s 45 -c 1 -o gpurun_out/r2aa_kwave5_c4 -f python bench.py --config c4 $B > gpurun_out/r2aa_ncu.log 2>&1
ls -la gpurun_out/r2aa*
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "torchrun n1 rc=$?"
