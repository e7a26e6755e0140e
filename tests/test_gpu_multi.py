"""Multi-GPU tests (need >= 2 GPUs on the box; skipped otherwise): the data-parallel regression step with the gradient
all-reduce fused into the optimiser kernel over NVLink peer memory (picard/solution_jac.py:167-213)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_two_rank_data_parallel_fit_matches_single_gpu_and_keeps_replicas_identical(tmp_path):
    out = tmp_path / "dp.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "tests", "dp_worker.py"), str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=540)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    assert all(x["replicas_bitwise_identical"] for x in res), res
    r0 = res[0]
    print("data-parallel fit:", r0)
    # same global batches, same steps; the only difference is the summation order of the batch mean (two halves)
    assert r0["rel_diff_vs_single_gpu"] < 1e-4, r0
