"""GPU, >= 2 devices: the multi-GPU path through torchrun (skipped on one GPU; the host-side logic of the
same path is covered on CPU by tests/test_distributed_cpu.py over gloo).

* one-sided NVLink transport vs the expected values, bit for bit (tools/check_p2p.py),
* distributed AMG-PCG vs Jacobi-PCG on 2 ranks, partitioned + agglomerated levels (tools/debug_dist_amg.py),
* P2 state / adjoint pipeline on a partition vs the one-GPU result (tools/check_dist_p2.py),
* distance-based mu and re-extension on a partition vs the one-GPU result (tools/check_dist_features.py).
"""

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, *args, port=29611, nproc=2, timeout=240):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script), *map(str, args)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    return out.stdout + out.stderr


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


@needs2
def test_one_sided_transport():
    out = _torchrun("check_p2p.py", 12, port=29611)
    assert out.count("P2P_OK") == 2 and "P2P_FAILED" not in out, out[-2000:]


@needs2
@pytest.mark.parametrize("args", [(24,), (48, 500)])
def test_distributed_amg(args):
    out = _torchrun("debug_dist_amg.py", *args, port=29612)
    assert "DIST_AMG_OK" in out, out[-2000:]


@needs2
@pytest.mark.parametrize("pc", ["jacobi", "hypre"])
def test_distributed_p2_pipeline(pc):
    out = _torchrun("check_dist_p2.py", 8, pc, port=29613)
    assert "DIST_P2_OK" in out, out[-2000:]


@needs2
def test_distributed_round2_features():
    out = _torchrun("check_dist_features.py", 10, port=29614)
    assert "DIST_FEATURES_OK" in out, out[-2000:]
