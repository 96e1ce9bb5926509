timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/perf_probe.py c2 c3 c4 2>&1 | tee gpurun_out/perf12.log
RT_B200_SCHED=waves timeout 300 python tools/perf_probe.py c3 2>&1 | tee -a gpurun_out/perf12.log
