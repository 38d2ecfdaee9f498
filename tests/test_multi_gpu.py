"""Real multi-GPU run of both sharding modes over NCCL (tools/multi_gpu_check.py under torchrun).
Skipped on boxes with a single GPU — there the same logic is covered by the world-size-2 gloo test
(tests/test_sharding_gloo.py) and the one-device shard emulation (test_gpu_parity.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_and_point_sharding_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["frames_identical"] and rep["points_keys_identical"] and rep["points_image_identical"]
    assert rep["fused_keys_identical"] and rep["fused_image_identical"] and rep["zmerge_nccl_identical"]
