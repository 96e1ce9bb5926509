timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "--- sah collapse"; python tools/perf_probe.py c3 c2 c4 2>&1 | cut -c1-420
echo "--- even collapse"; RT_B200_COLLAPSE=even python tools/perf_probe.py c3 c4 2>&1 | cut -c1-420
