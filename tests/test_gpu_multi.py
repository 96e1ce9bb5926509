"""Multi-GPU gather paths on real GPUs (skipped on a single-GPU box; the host-side sharding logic is
covered on CPU by tests/test_sharding_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_p2p_landing_and_nccl_gather_assemble_the_full_frame(gpu_present):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 4 if n >= 4 else 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "p2p_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "first p2p landing == full frame: True" in r.stdout and "first p2p landing == full frame: False" not in r.stdout
    assert "p2p landing == full frame: True" in r.stdout and "nccl gather == full frame: True" in r.stdout
    # back-pressure: four frames pushed back to back into ONE landing buffer whose consumer is slow -- no snapshot shows a later frame
    assert r.stdout.count("back-pressure: snapshot of frame") == 4 and "back-pressure: snapshot of frame" in r.stdout
    assert "full frame of camera 0: True" in r.stdout and "full frame of camera 3: True" in r.stdout and "of camera 1: False" not in r.stdout and "of camera 2: False" not in r.stdout
