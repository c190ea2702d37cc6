"""CPU: host logic of the scan - grid construction (against the reference's formula), rank
sharding, and the world_size-2 gather over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from molearn_b200.analysis import MolearnAnalysis, shard_range


class TinyAE(torch.nn.Module):
    """encode/decode with the reference's shapes ([B,N,3] <-> [B,2]; decode pads to M >= N)."""

    def __init__(self, n=20, m=32):
        super().__init__()
        g = torch.Generator().manual_seed(3)
        self.enc = torch.nn.Parameter(torch.randn(n * 3, 2, generator=g) * 0.1)
        self.dec = torch.nn.Parameter(torch.randn(2, m * 3, generator=g))
        self.m = m

    def encode(self, x):
        return x.reshape(x.shape[0], -1) @ self.enc

    def decode(self, z):
        return torch.tanh(z.reshape(z.shape[0], 2) @ self.dec).reshape(-1, 3, self.m).permute(0, 2, 1)


class Data:
    def __init__(self, n_frames=6, n=20):
        g = torch.Generator().manual_seed(5)
        self.dataset = torch.randn(n_frames, n, 3, generator=g)
        self.std, self.mean, self.standardize = 2.5, 0.3, True
        self.atoms = ["CA"]

    def get_atominfo(self):
        return np.array([("CA", "ALA", i) for i in range(self.dataset.shape[1])], dtype=object)


def make(batch_size=7):
    ma = MolearnAnalysis(batch_size=batch_size)
    ma.set_network(TinyAE())
    ma.set_dataset("train", Data())
    return ma


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 64, 65536, 16384):
        for world in (1, 2, 3, 8):
            got, per = [], None
            for r in range(world):
                lo, hi, per = shard_range(total, r, world)
                assert hi - lo <= per
                got += list(range(lo, hi))
            assert got == list(range(total))


def test_registry_views():
    """`datasets` / `encoded` read-only views (analyser.py:1195-1201) that the reference's GUI and plot code read."""
    ma = make()
    assert list(ma.datasets) == ["train"] and ma.datasets["train"].shape == ma.get_dataset("train").shape
    ma.get_encoded("train")
    ma.setup_grid(samples=4)
    assert set(ma.encoded) == {"train", "grid"}
    ma.encoded.pop("grid")                                # a copy: the registry itself is untouched
    assert "grid" in ma.encoded


def test_setup_grid_matches_reference_formula():
    """analyser.py:568-594: linspace over padded bounds, meshgrid (row = y), reshape(-1,1,2)."""
    ma = make()
    ma.setup_grid(samples=9, padding=0.1)
    z = ma.get_encoded("train").numpy()
    x0, x1, y0, y1 = z[:, 0].min(), z[:, 0].max(), z[:, 1].min(), z[:, 1].max()
    bx, by = (x1 - x0) * 0.1, (y1 - y0) * 0.1
    xs, ys = np.linspace(x0 - bx, x1 + bx, 9), np.linspace(y0 - by, y1 + by, 9)
    assert np.allclose(ma.xvals, xs) and np.allclose(ma.yvals, ys)
    grid = ma._encoded["grid"]
    assert grid.shape == (81, 1, 2) and grid.dtype == torch.float32
    assert np.allclose(grid[9 * 4 + 2, 0].numpy(), [xs[2], ys[4]], atol=1e-6)      # row-major: y outer, x inner
    ma.setup_grid(samples=4, bounds=(-1, 1, -2, 2), padding=0.5)
    assert np.allclose(ma.xvals, np.linspace(-2, 2, 4)) and np.allclose(ma.yvals, np.linspace(-4, 4, 4))


def test_scan_error_semantics_and_cache():
    ma = make()
    ma.setup_grid(samples=5)
    rmsd, drift, xv, yv = ma.scan_error()
    assert rmsd.shape == (5, 5) and drift.shape == (5, 5)
    assert set(ma.surfaces) == {"Network_RMSD", "Network_z_drift"}
    net = ma.network
    z = ma._encoded["grid"].reshape(25, 2)
    with torch.no_grad():
        d1 = net.decode(z)[:, :20]
        z2 = net.encode(d1)
        d2 = net.decode(z2)[:, :20]
    want = (((d1 - d2) * 2.5) ** 2).sum(2).mean(1).sqrt().reshape(5, 5).numpy()
    assert np.allclose(rmsd, want, rtol=1e-5)
    assert np.allclose(drift, ((z - z2) ** 2).mean(1).sqrt().reshape(5, 5).numpy(), rtol=1e-5)
    assert ma.scan_error()[0] is rmsd                                    # cached surface
    with pytest.raises(ValueError):
        make().scan_error()                                              # no grid yet


def test_get_decoded_and_custom_scan():
    ma = make(batch_size=4)
    ma.setup_grid(samples=3)
    dec = ma.get_decoded("grid")
    assert dec.shape == (9, 20, 3)
    surf, _, _ = ma.scan_custom(lambda s, k: float(np.abs(s).max()) * k, [2.0], "absmax")
    assert surf.shape == (3, 3) and np.isclose(surf[0, 0], float((dec[0] * 2.5).abs().max()) * 2.0)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ma = make(batch_size=3)
    ma.setup_grid(samples=5)
    rmsd, drift, _, _ = ma.scan_error()
    if rank == 0:
        np.save(out, np.stack([rmsd, drift]))
        assert ma.scan_stats["points_this_rank"] == 13 and ma.scan_stats["world"] == 2
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scan_equals_single_rank(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r0.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ma = make(batch_size=3)
    ma.setup_grid(samples=5)
    rmsd, drift, _, _ = ma.scan_error()
    got = np.load(out)
    assert np.array_equal(got[0], rmsd) and np.array_equal(got[1], drift)


def test_backbone_angles_and_dihedrals_on_md_frames():
    """get_angles / get_dihedrals on real MurD frames: textbook backbone geometry (omega ~ 180 deg, N-CA-C ~ 111 deg)."""
    from conftest import load_golden
    g = load_golden("murd_input")
    keep = np.isin(g["names"], ("N", "CA", "CB", "C", "O"))
    ma = MolearnAnalysis()
    ma._atominfo = np.array(list(zip(g["names"][keep], g["resnames"][keep], g["resids"][keep])), dtype=object)
    x = g["coords"][:2, keep]
    dih = ma._get_dihedrals(x)
    ang = ma._get_angles(x)
    assert dih["Phi"].shape == (2, 436) and dih["Improper Torsion"].shape == (2, 397)      # 437 residues, 40 GLY
    assert np.median(np.abs(np.degrees(dih["Omega"]))) > 170                                # trans peptide bonds
    assert abs(np.degrees(ang["N-CA-C"]).mean() - 111) < 3 and abs(np.degrees(ang["CA-C-O"]).mean() - 120.5) < 3
    assert np.all(np.degrees(dih["Improper Torsion"]) > 0) or np.all(np.degrees(dih["Improper Torsion"]) < 0)   # one chirality
