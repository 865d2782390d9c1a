"""Multi-GPU parity over NCCL (run with `-m gpu` on a box with >= 2 B200s; skipped on one GPU).

`tools/check_sharded.py` is launched with torchrun, one rank per GPU: the z-slab sharded prediction (slab input with
exchanged uint8 strips, and replicated input; probabilities exchanged in one and in several ragged chunks; 2 and 4
classes) must equal the single-GPU prediction bit for bit.  The log is kept under `gpurun_out/` when that exists."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_prediction_is_bit_identical_over_nccl(built_library, world):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    log_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log_dir):
        with open(os.path.join(log_dir, f"sharded_check_{world}gpu.txt"), "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_CHECK PASS" in r.stdout
    assert r.stdout.count("bit-identical on every rank = True") >= 4
