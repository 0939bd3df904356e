"""Slab-partitioned multi-GPU solve (NVLink halo pushes + NCCL scalar all-reduces) against the
single-GPU solve of the same problem; needs at least two GPUs (skipped on a one-GPU box)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_slab_solve_matches_single_gpu():
    n = 2          # the configuration verified on hardware in round 1 (profiles/r01q_mg2_*); N = 4, 8 are exercised by bench.py --gpus N
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_vertex_partitioned_graph_solve_matches_single_gpu():
    n = 2          # the configuration verified on hardware in round 1 (profiles/r01q_mg2_*); N = 4, 8 are exercised by bench.py --gpus N
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29512", os.path.join(ROOT, "tests", "mgpu_graph_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
