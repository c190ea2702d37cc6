"""GPU: latent-grid scan and physics trainer through the public classes, checked against the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import energy_oracle as eo

pytestmark = pytest.mark.gpu
ATOMS = ("N", "CA", "CB", "C", "O")


@pytest.fixture(scope="module")
def murd_data():
    from molearn_b200.data import CoordinateSet
    d = load_golden("murd_input")
    keep = np.isin(d["names"], ATOMS)
    info = np.empty((int(keep.sum()), 3), dtype=object)
    info[:, 0], info[:, 1], info[:, 2] = list(d["names"][keep]), list(d["resnames"][keep]), [int(r) for r in d["resids"][keep]]
    return CoordinateSet(info, np.ascontiguousarray(d["coords"][:, keep])).prepare_dataset()


@pytest.fixture(scope="module")
def analysis(murd_data):
    from molearn_b200.analysis import MolearnAnalysis
    from molearn_b200.models import AutoEncoder
    torch.manual_seed(0)
    net = AutoEncoder(out_points=murd_data.dataset.shape[1]).to("cuda:0")
    # torch's default convolution math (TF32 on), as bench.py's scans run.  One decode batch = the whole 6x6 grid, so
    # the test's own decode of the grid reproduces the structures the scan scored (same kernels, same batch); the
    # kernels are then checked on exactly those structures.
    ma = MolearnAnalysis(batch_size=36)
    ma.set_network(net)
    ma.set_dataset("train", murd_data)
    ma.setup_grid(samples=6, padding=0.1)
    return ma


def decoded_angstrom(ma):
    with torch.no_grad():
        z = ma._encoded["grid"].reshape(-1, 2).to(ma.device)
        return (ma._decode(z) * ma.stdval + ma.meanval).double().cpu().numpy()       # the scans' own decode path


def test_scan_physics_vs_oracle(analysis):
    ma = analysis
    total, xv, yv = ma.scan_physics()
    assert total.shape == (6, 6) and len(xv) == 6 and len(yv) == 6
    topo = load_golden("topo_bbcb")
    dec = decoded_angstrom(ma)
    for p in range(36):                                                  # every grid point
        want, _ = eo.energy_all(dec[p], topo, want_grad=False)
        got = [ma.surfaces[f"Physics_{n}"].reshape(-1)[p] for n in ("bond", "angle", "torsion", "nb")]
        for k in range(4):
            assert abs(got[k] - want[k]) < 2e-5 * abs(want[k]), (p, k, got[k], want[k])   # the oracle sees the fp32 structure of a second, identical decode
        assert abs(total.reshape(-1)[p] - want.sum()) < 2e-5 * abs(want.sum())


def test_scan_physics_fp32_decode_vs_oracle(analysis):
    """The same check with cuDNN's TF32 convolution math off (fp32 decode), all 36 points."""
    ma = analysis
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for k in list(ma.surfaces):
            if k.startswith("Physics"):
                del ma.surfaces[k]
        total, _, _ = ma.scan_physics()
        topo = load_golden("topo_bbcb")
        dec = decoded_angstrom(ma)
        for p in range(36):
            want, _ = eo.energy_all(dec[p], topo, want_grad=False)
            got = [ma.surfaces[f"Physics_{n}"].reshape(-1)[p] for n in ("bond", "angle", "torsion", "nb")]
            for k in range(4):
                assert abs(got[k] - want[k]) < 2e-5 * abs(want[k]), (p, k, got[k], want[k])
    finally:
        torch.backends.cudnn.allow_tf32 = old
        for k in list(ma.surfaces):
            if k.startswith("Physics"):
                del ma.surfaces[k]
        ma.scan_physics()                                               # restore the TF32 surfaces the later tests compare with


def test_scan_scores_at_matches_scan(analysis, murd_data):
    """Re-scoring chosen grid points on one rank (scan_scores_at: bench.py's sharded-vs-single check) gives the rows of
    the full scan."""
    ma = analysis
    tg = torch.from_numpy(murd_data.coords[:2])
    rows = ma.scan_scores(targets=tg)
    # all 36 points in another order: the decode batch keeps its shape, so cuDNN runs the same kernels and every
    # structure is reproduced bit for bit (a batch of another size may use another TF32 algorithm, and the energies of
    # an untrained decoder's collapsed structures move by per cent under TF32-sized coordinate noise)
    idx = torch.randperm(36, generator=torch.Generator().manual_seed(3))
    again = ma.scan_scores_at(idx, targets=tg)
    assert again.shape == (36, 6)
    assert torch.equal(again, rows[idx]), (again - rows[idx]).abs().max()


def test_scan_error_from_target(analysis, murd_data):
    ma = analysis
    surf, _, _ = ma.scan_error_from_target("train", index=3, align=False)
    dec = decoded_angstrom(ma)
    tgt = murd_data.coords[3].astype(np.float64)
    want = np.sqrt(((dec - tgt) ** 2).sum(2).mean(1)).reshape(6, 6)
    assert np.allclose(surf, want, rtol=2e-5)
    assert "RMSD_from_train_index_3" in ma.surfaces
    del ma.surfaces["RMSD_from_train_index_3"]
    aligned, _, _ = ma.scan_error_from_target("train", index=3)         # align=True is the reference default
    assert np.all(aligned <= surf * (1 + 1e-5))                          # superposition can only lower the RMSD
    del ma.surfaces["RMSD_from_train_index_3"]
    err = ma.get_error("train", align=True)
    assert err.shape == (16,) and np.all(err <= ma.get_error("train", align=False) * (1 + 1e-5))
    with pytest.raises(Exception):
        ma.scan_error_from_target("train")                               # 16 frames, no index


def test_scan_scores_and_error(analysis, murd_data):
    ma = analysis
    tg = torch.from_numpy(murd_data.coords[:4])
    rows = ma.scan_scores(targets=tg)
    assert rows.shape == (36, 8)
    assert np.allclose(rows[:, 0].numpy().reshape(6, 6), ma.surfaces["Physics_bond"], rtol=1e-5)
    ma.surfaces.pop("RMSD_from_train_index_3", None)
    assert np.allclose(rows[:, 7].numpy().reshape(6, 6), ma.scan_error_from_target("train", index=3, align=False)[0], rtol=1e-5)
    rmsd, drift, _, _ = ma.scan_error()
    assert rmsd.shape == (6, 6) and np.isfinite(rmsd).all() and np.isfinite(drift).all()


def test_physics_trainer_step(murd_data):
    from molearn_b200.models import AutoEncoder
    from molearn_b200.trainers import Torch_Physics_Trainer
    torch.manual_seed(0)
    tr = Torch_Physics_Trainer(device="cuda:0")
    tr.set_data(murd_data)
    tr.set_autoencoder(AutoEncoder, out_points=murd_data.dataset.shape[1])
    tr.prepare_optimiser()
    tr.prepare_physics(physics_scaling_factor=0.1)
    batch = murd_data.dataset[:8].to("cuda:0")
    before = [p.detach().clone() for p in tr.autoencoder.decoder.parameters()]
    res = tr.training_iteration(batch)
    for k in ("physics_loss", "bond_energy", "angle_energy", "torsion_energy", "mse_loss", "loss"):
        assert k in res and torch.isfinite(res[k]).all(), k
    assert any(not torch.equal(a, b) for a, b in zip(before, tr.autoencoder.decoder.parameters()))
    # the physics terms are the oracle's bonded energies of the generated structures
    gen = (tr._internal["generated"].detach() * tr.std + tr.mean).double().cpu().numpy()
    topo = load_golden("topo_bbcb")
    want = np.mean([eo.bonded(g, topo, want_grad=False)[0] for g in gen], axis=0)
    got = [float(res[k]) for k in ("bond_energy", "angle_energy", "torsion_energy")]
    for k in range(3):
        assert abs(got[k] - want[k]) < 1e-5 * want[k]
    # and the physics gradient reaches the decoder: d(physics)/d(decoder params) != 0
    tr.autoencoder.zero_grad()
    out = tr.common_step(batch)
    out.update(tr.common_physics_step(batch, tr._internal["encoded"]))
    out["physics_loss"].backward()
    gnorm = sum(float(p.grad.abs().sum()) for p in tr.autoencoder.decoder.parameters() if p.grad is not None)
    assert gnorm > 0 and np.isfinite(gnorm)
    v = tr.valid_step(batch)
    assert torch.isfinite(v["loss"])


def test_physics_energy_module(murd_data):
    """[B,n,3] standardised in -> [B] energies out, differentiable (the openmm_energy call contract)."""
    from molearn_b200.loss_functions import physics_energy
    mod = physics_energy(murd_data.get_atominfo(), std=murd_data.std, mean=murd_data.mean, device="cuda:0",
                         clamp={"min": -1e3, "max": 1e3})
    x = murd_data.dataset[:3].to("cuda:0").requires_grad_(True)
    e = mod(x)
    assert e.shape == (3,)
    e.nanmean().backward()
    assert torch.isfinite(x.grad).all() and float(x.grad.abs().max()) <= 1e3 * murd_data.std / 3 + 1e-3
    topo = load_golden("topo_bbcb")
    want = eo.energy_all(murd_data.coords[0].astype(np.float64), topo, want_grad=False)[0].sum()
    assert abs(float(e[0]) - want) < 2e-5 * abs(want)


def test_geometry_scans_vs_numpy(analysis, murd_data):
    """scan_bondlength / scan_ca_chirality (device, fused) against the reference's NumPy formulas
    (analyser.py:831-861) applied to the decoded structures."""
    from molearn_b200.analysis import backbone_indices
    ma = analysis
    ma.scan_bondlength()
    ma.scan_ca_chirality()
    dec = decoded_angstrom(ma) - ma.meanval                                # the reference's scan uses decoded*std
    bonds, quads = backbone_indices(murd_data.get_atominfo())
    assert set(bonds) == {"N-CA", "CA-C", "C-N", "CA-CB"} and len(bonds["C-N"]) == len(bonds["N-CA"]) - 1
    for k, v in bonds.items():
        L = np.array([np.linalg.norm(dec[:, i, :] - dec[:, j, :], axis=1) for i, j in v])   # [n_bonds, points]
        assert np.allclose(ma.surfaces[k].reshape(-1), L.mean(axis=0), rtol=1e-4, atol=1e-5), k
        assert np.allclose(ma.surfaces[k + "_std"].reshape(-1), L.std(axis=0), rtol=1e-3, atol=1e-4), k
    q = np.asarray(quads)
    n, ca, c, cb = (dec[:, q[:, m], :] for m in range(4))
    trip = np.einsum("bij,bij->bi", np.cross(n - ca, c - ca), cb - ca)
    margin = np.abs(trip).min(axis=1) > 1e-6                                # ignore points with a borderline centre
    assert np.array_equal(ma.surfaces["Chirality"].reshape(-1)[margin], (trip < 0).sum(axis=1)[margin])
    inv = ma.get_inversions("train")
    assert inv["dataset_inversions"].shape == (16,) and (inv["dataset_inversions"] == 0).all()   # MD frames: no D-residues
    bl = ma.get_bondlengths("train")
    assert abs(bl["dataset_bondlen"]["N-CA"].mean() - 1.46) < 0.03


def test_fused_decoder_matches_plain_decode_gpu():
    """models.FusedDecoder (BatchNorm folded, cuDNN fused conv+bias+relu, channels-last) against network.decode:
    fp32 convolution math -> 1e-4; cuDNN's default TF32 math -> both sit within TF32 noise of each other."""
    from molearn_b200.models import AutoEncoder, FusedDecoder
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = AutoEncoder(out_points=2145).to(dev).eval()
    g = torch.Generator().manual_seed(1)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(0.3 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.bias.data.copy_(0.2 * torch.randn(m.num_features, generator=g))
    fused = FusedDecoder(net.decoder)
    z = torch.rand(16, 2, device=dev) * 2 - 1
    old = torch.backends.cudnn.allow_tf32
    try:
        for tf32, tol in ((False, 1e-4), (True, 2e-2)):
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                want = net.decode(z)
                got = fused(z)
            assert got.shape == want.shape
            err = float((got - want).norm() / want.norm())
            assert err < tol, (tf32, err)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert fused.fused_ok, "cuDNN fused conv+bias+relu was not available: the decoder ran the unfused fallback"


def test_internal_coordinates_on_device_vs_numpy(analysis, murd_data):
    """get_bondlengths / get_angles / get_dihedrals / get_inversions evaluate their index groups on the device
    (mlp_internal_coords_eval; nothing of size [F, N, 3] is staged on the host).  Checked against the reference's NumPy
    formulas (analyser.py:791-861, the dihedral in its standard form) applied to the same structures, for the MD frames
    and for their decoded counterparts."""
    ma = analysis
    enc = ma.get_encoded("train")
    # the device works on the structures as stored (standardised units; distances are rescaled by std, angles need no
    # rescaling), so the NumPy side is given the same numbers: an untrained decoder's collapsed output (atoms 1e-3 A
    # apart) does not survive a *std + mean round trip in fp32
    with torch.no_grad():
        dec = ma._decode(enc.to(ma.device)).double().cpu().numpy() * ma.stdval
    ds = ma.get_dataset("train").double().numpy() * ma.stdval
    dih, ang, bl = ma.get_dihedrals("train"), ma.get_angles("train"), ma.get_bondlengths("train")
    for which, x in (("dataset", ds), ("decoded", dec)):
        want_d, want_a = ma._get_dihedrals(x), ma._get_angles(x)
        assert set(dih[f"{which}_dihedrals"]) == set(want_d) == {"Phi", "Psi", "Omega", "Improper Torsion"}
        for k, w in want_d.items():
            got = dih[f"{which}_dihedrals"][k]
            assert got.shape == w.shape == (16, len(w[0]))
            d = np.abs(np.angle(np.exp(1j * (got - w))))                   # compare on the circle
            assert d.max() < 2e-5, (which, k, d.max())
        assert set(ang[f"{which}_angles"]) == set(want_a)
        for k, w in want_a.items():
            got = ang[f"{which}_angles"][k]
            assert got.shape == w.shape and np.abs(got - w).max() < 2e-5, (which, k)
        from molearn_b200.analysis import backbone_indices
        bonds, _ = backbone_indices(murd_data.get_atominfo())
        for k, v in bonds.items():
            w = np.array([np.linalg.norm(x[:, i, :] - x[:, j, :], axis=1) for i, j in v])
            got = bl[f"{which}_bondlen"][k]
            assert got.shape == w.shape and np.allclose(got, w, rtol=2e-5, atol=2e-5), (which, k)
    assert abs(np.degrees(ang["dataset_angles"]["N-CA-C"]).mean() - 111) < 3
    # a grid key has no dataset side
    only = ma.get_angles("grid")
    assert set(only) == {"decoded_angles"} and only["decoded_angles"]["N-CA-C"].shape == (36, 437)
