"""The multi-PROCESS slab path — one process per GPU, CUDA-IPC peer copies with device-side stamps (what bench.py --gpus N times)
and NCCL send/recv — against a single context: positions, velocities and bin ids bit for bit, pair sums <= 1e-5, particles
conserved, migration exercised.  Needs at least two GPUs on the box (skipped otherwise); launched through torch.distributed.run
exactly like the bench."""
import json
import pathlib
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_slab_processes_match_single_context(exchange):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if ngpu >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tools" / "slab_parity_mp.py"), "--grid", "64", "--steps", "6", "--exchange", exchange]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == world
    for k in ("symmetric_kernel", "gather_kernel"):
        r = out[k]
        assert r["conserved"], r
        assert r["bit_exact"] and r["cells_equal_frac"] == 1.0, r
        assert r["max_rel_l2"] <= 1e-5, r
        assert r["migrated"] > 0, "the drift was meant to carry particles across the slab faces"
