timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -1
python tools/perf_probe.py c3 2>&1 | cut -c1-200
python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-150
python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-150
