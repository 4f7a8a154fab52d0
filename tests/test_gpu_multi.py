"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): the real N > 1 path -- spp-sharded render on every rank, one
NCCL reduce of the device-resident SUM buffers -- must reproduce the single-GPU image to fp32 summation tolerance.
(The host logic alone is covered on CPU by tests/test_distributed_gloo.py.)"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_sharded_render_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    d = out["determinism"]
    assert out["world"] == 2 and d["samples_ok"]
    assert d["max_rel_diff"] < 5e-6
    assert out["c4"]["gpaths_per_s"] > 1.0
