"""kr_average_by_duration on the device against fixtures from the live reference function (utils/lengths.py:156-208).
(The kernel body is also verified bit-identical by host emulation, tests/test_lengths_emu_cpu.py.)"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_average_by_duration_bit_identical_to_reference_fixtures():
    from kokoro_ruslan_b200.lengths import average_by_duration
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "average_by_duration.npz"))
    for k in range(4):
        v, d, m = (torch.from_numpy(f[f"{n}{k}"]).cuda() for n in "vdm")
        assert torch.equal(average_by_duration(v, d).cpu(), torch.from_numpy(f[f"a{k}"])), k
        assert torch.equal(average_by_duration(v, d, m).cpu(), torch.from_numpy(f[f"am{k}"])), k
