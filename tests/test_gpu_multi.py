"""GPU suite: sharded execution on >= 2 B200s of one box (skipped on a single-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_sharded_plymouth_matches_oracle_on_all_visible_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tests" / "multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(out.stdout[-4000:], out.stderr[-4000:])
    assert out.returncode == 0
    assert f"MULTI_GPU_OK world={n}" in out.stdout


@pytest.mark.gpu
def test_pipelined_executions_stress_single_gpu():
    """scripts/stress_pipeline.py: random bursts of back-to-back (pipelined) executions, a second query interleaved,
    option flips -- every fetched result equals the oracle's."""
    import os
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "stress_pipeline.py")], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, ITERS="80"))
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert out.returncode == 0 and "STRESS_OK world=1" in out.stdout


@pytest.mark.gpu
def test_pipelined_executions_stress_all_visible_gpus():
    import os
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29547", str(ROOT / "scripts" / "stress_pipeline.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, ITERS="80"))
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert out.returncode == 0 and f"STRESS_OK world={n}" in out.stdout
