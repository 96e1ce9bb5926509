mkdir -p gpurun_out
cd /root/repo
( timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_timed_path.py -x -q -m gpu -k "test_forced_scheduler and (waves-bin or waves-defer) and (c4 or t_mesh)" 2>&1 | tail -15 ) > gpurun_out/r2y_memcheck.txt; echo "memcheck rc=$?"; tail -8 gpurun_out/r2y_memcheck.txt
( timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_timed_path.py -x -q -m gpu -k "test_forced_scheduler and waves-defer and t_mesh" 2>&1 | tail -15 ) > gpurun_out/r2y_racecheck.txt; tail -8 gpurun_out/r2y_racecheck.txt
( timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 ) > gpurun_out/r2y_smoke_memcheck.txt; tail -4 gpurun_out/r2y_smoke_memcheck.txt
