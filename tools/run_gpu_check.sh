timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for l in 2 1 3 4; do echo "--- leaf $l"; RT_B200_LEAF_SIZE=$l python tools/perf_probe.py c3 c4 2>&1 | cut -c1-400; done
python tools/perf_probe.py c2 2>&1 | cut -c1-400
