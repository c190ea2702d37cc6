import numpy as np
import pytest
import torch

from oracle import energy_oracle as eo

pytestmark = pytest.mark.gpu


def test_rmsd_kernel_vs_oracle():
    from molearn_b200.analysis.reducers import rmsd_to_targets
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    x = rng.normal(size=(7, 301, 3)).astype(np.float32)
    t = rng.normal(size=(3, 301, 3)).astype(np.float32) * 4 + 1
    out = rmsd_to_targets(torch.from_numpy(x).to(dev), torch.from_numpy(t).to(dev), scale=2.5, shift=-0.7).cpu().numpy()
    for b in range(7):
        for k in range(3):
            want = eo.rmsd(x[b].astype(np.float64) * 2.5 - 0.7, t[k])
            assert abs(out[b, k] - want) < 1e-5 * want
