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


def kabsch_rmsd(a, b):
    """NumPy reference: optimal-superposition RMSD via SVD (Kabsch 1976)."""
    a = a - a.mean(0)
    b = b - b.mean(0)
    u, s, vt = np.linalg.svd(a.T @ b)
    if np.linalg.det(u @ vt) < 0:
        s[-1] = -s[-1]
    return np.sqrt(max((a * a).sum() + (b * b).sum() - 2 * s.sum(), 0.0) / len(a))


def test_aligned_rmsd_vs_numpy_kabsch():
    """biobox is absent (parity unpinned for the aligned variant): checked against a NumPy Kabsch."""
    from molearn_b200.analysis.reducers import rmsd_to_targets
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    base = rng.normal(size=(400, 3)) * 15 + 40
    xs, ts = [], []
    for i in range(5):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        xs.append((base + rng.normal(size=base.shape) * (0.02 + 0.5 * i)) @ q.T + rng.normal(size=3) * 30)
    for i in range(3):
        ts.append(base + rng.normal(size=base.shape) * 0.3 * i)
    ts.append(base * np.array([1, 1, -1.0]))                           # mirror image: needs the proper-rotation fix
    x = np.stack(xs).astype(np.float32)
    t = np.stack(ts).astype(np.float32)
    out = rmsd_to_targets(torch.from_numpy(x).to(dev), torch.from_numpy(t).to(dev), align=True).cpu().numpy()
    for b in range(len(xs)):
        for k in range(len(ts)):
            want = kabsch_rmsd(x[b].astype(np.float64), t[k].astype(np.float64))
            assert abs(out[b, k] - want) < 1e-4 * max(want, 1e-2), (b, k, out[b, k], want)
    # standardised input + scale/shift, strided slice
    pad = torch.zeros(5, 450, 3, device=dev)
    pad[:, :400] = torch.from_numpy((x - 40.0) / 15.0).to(dev)
    out2 = rmsd_to_targets(pad[:, :400], torch.from_numpy(t).to(dev), scale=15.0, shift=40.0, align=True).cpu().numpy()
    assert np.allclose(out2, out, rtol=1e-3, atol=1e-3)


def test_paired_rmsd_and_input_checks():
    """MLP_RMSD_PAIRED (get_error: structure b against target b, no B x B matrix), K > 8 targets per structure (several
    blocks of targets), and the dtype / device / shape checks of the Python face."""
    from molearn_b200.analysis.reducers import rmsd_to_targets
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(8)
    x = torch.from_numpy(rng.normal(size=(19, 257, 3)).astype(np.float32) * 7).to(dev)
    t = torch.from_numpy(rng.normal(size=(19, 257, 3)).astype(np.float32) * 7 + 1).to(dev)
    for align in (False, True):
        full = rmsd_to_targets(x, t, align=align)
        pair = rmsd_to_targets(x, t, align=align, paired=True)
        assert full.shape == (19, 19) and pair.shape == (19,)
        assert torch.equal(pair, full.diagonal())
    want = np.sqrt(((x.double().cpu().numpy()[:, None] - t.double().cpu().numpy()[None]) ** 2).sum(-1).mean(-1))
    assert np.allclose(rmsd_to_targets(x, t).cpu().numpy(), want, rtol=1e-5)
    with pytest.raises(ValueError):
        rmsd_to_targets(x, t[:5], paired=True)
    with pytest.raises(TypeError):
        rmsd_to_targets(x.double(), t)
    with pytest.raises(ValueError):
        rmsd_to_targets(x, t[:, :100])
    with pytest.raises(RuntimeError):
        rmsd_to_targets(x.cpu(), t)
