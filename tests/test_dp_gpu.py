"""Config 3 of BASELINE.json (data-parallel training) as a device test: tools/dp_check.py under torchrun on 2 GPUs —
replicas bit-identical (also when one rank alone draws a long utterance and tightens its clip: the clip is a global MIN
inside the collective kernel), losses and the accumulated update equal to the single-process accumulation window, eager
and CUDA-graph steps.  Skipped on boxes with fewer than 2 devices (the driver's GPU test tier has 1)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("multicast", ["0", "1"])
def test_two_rank_training_matches_single_process_window(multicast):
    """multicast = 1 forces the NVSwitch multicast path at 2 GPUs (the default there is peer loads / stores), which is
    also the path that overlaps the early gradient ranges with the backward tail (two launches from the comm stream)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=540,
                       env={**os.environ, "KR_MULTICAST": multicast})
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
