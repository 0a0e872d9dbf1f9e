"""Multi-GPU correctness of the fused data-parallel tails, inside the driver-run suite: each test launches a 2-rank
torchrun job (one process per GPU, NCCL + NVLink peer memory) and checks its verdict.  Skipped on boxes with < 2 GPUs.

  dp_fused_sgd_worker.py   salun_dp_masked_sgd_step (reduce-scatter + masked SGD + all-gather in one kernel) vs NCCL
                           all-reduce + salun_masked_sgd_step: weights within 2e-6, masked-out coordinates exact,
                           replicas bit-identical
  dp_fused_adam_worker.py  salun_dp_grad_reduce_sumsq + salun_dp_masked_adam_step vs NCCL all-reduce + fused clip/Adam,
                           then one DDPM saliency_unlearn iteration through the runner on both paths
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(worker, nproc=2, timeout=420):
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "dist", worker)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_fused_dp_masked_sgd_two_ranks():
    out = _run("dp_fused_sgd_worker.py")
    assert "replicas identical: True" in out and "masked-out exact: True" in out


def test_graphed_dp_step_equals_eager_two_ranks():
    """dp_graph_worker.py: the whole data-parallel step, symmetric-memory barriers and the fused exchange kernel included,
    replayed from a CUDA graph"""
    out = _run("dp_graph_worker.py")
    assert "graphed DP step equals eager: True" in out and "replicas identical: True" in out


def test_sync_bn_sharded_step_equals_full_batch_step():
    """dp_syncbn_worker.py: sync-BN over NVLink peer memory -- a batch sharded over 2 ranks reproduces the single-process
    step of the concatenated batch (weights, running statistics); per-shard statistics do not"""
    out = _run("dp_syncbn_worker.py")
    assert "replicas identical: True" in out


def test_fused_dp_masked_adam_two_ranks():
    out = _run("dp_fused_adam_worker.py")
    assert "replicas identical: True" in out and "norms ok: True" in out
