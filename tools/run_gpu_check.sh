timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "--- refill on"; python tools/perf_probe.py c3 2>&1 | cut -c1-330; RT_PIPE_SHARE="4" RT_PIPE_M="3" python tools/pipe_probe.py c3 1 2>&1 | tail -1
echo "--- refill off"; RT_B200_REFILL=0 python tools/perf_probe.py c3 2>&1 | cut -c1-330; RT_B200_REFILL=0 RT_PIPE_SHARE="4" RT_PIPE_M="3" python tools/pipe_probe.py c3 1 2>&1 | tail -1
ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_frame -s 4 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --pipelines 1 2>&1 | grep -E "ratio|inst_executed|duration|issue_active"
