for r in 64 256 1024 4096; do echo "--- RT_B200_RETIRE_RAYS=$r"
RT_B200_RETIRE_RAYS=$r RT_PIPE_SHARE="0" RT_PIPE_M="1 2 3" python tools/pipe_probe.py c3 1 2>&1 | tail -3
RT_B200_RETIRE_RAYS=$r RT_PIPE_SHARE="0" RT_PIPE_M="1 2 4" python tools/pipe_probe.py c3 8 2>&1 | tail -3
done
