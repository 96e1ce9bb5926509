timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
python tools/perf_probe.py c3 c2 c4 2>&1 | cut -c1-330
