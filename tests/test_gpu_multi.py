"""Partitioned multi-GPU solves (boundary layers pushed into the neighbours' memory over NVLink, PCG scalars
all-reduced inside the kernels over peer mailboxes) against the single-GPU solve of the same problem, at every
world size the box offers: 2 ranks have no interior rank, 4 and 8 do (both neighbours).  Skipped when fewer GPUs
are visible."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def torchrun(n, port, script, *args, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", script)] + list(args)
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=e)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    return r.stdout


@pytest.mark.parametrize("n", [2, 4, 8])
def test_slab_partitioned_solves_match_single_gpu(n):
    if NGPU < n:
        pytest.skip("needs %d GPUs" % n)
    out = torchrun(n, 29510 + n, "mgpu_check.py")
    assert "deterministic" in out


@pytest.mark.parametrize("n", [2, 4, 8])
def test_vertex_partitioned_graph_and_replicated_camera_solves_match_single_gpu(n):
    if NGPU < n:
        pytest.skip("needs %d GPUs" % n)
    torchrun(n, 29530 + n, "mgpu_graph_check.py")


@pytest.mark.skipif(NGPU < 2, reason="needs two GPUs")
def test_nccl_scalar_path_still_matches():
    """THALLO_B200_MG_NCCL=1: the comparison path (NCCL all-reduce of the PCG scalars, separate push / close kernels)."""
    torchrun(2, 29550, "mgpu_check.py", env={"THALLO_B200_MG_NCCL": "1"})
