set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python tools/gpu_check.py t_mixed t_ballplane t_mesh c4 2>&1 | tee gpurun_out/gpu_check2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tee gpurun_out/pytest_gpu.log | tail -30
