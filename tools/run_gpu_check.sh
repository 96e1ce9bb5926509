timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for o in 8 10 12; do echo "--- RT_B200_OCC=$o"; RT_B200_OCC=$o python tools/perf_probe.py c3 c2 c4 2>&1 | cut -c1-200; done
