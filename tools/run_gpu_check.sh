timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python tools/perf_probe.py c2 c3 c4 2>&1 | tee gpurun_out/perf10.log
