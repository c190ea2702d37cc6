"""GPU parity tests: the CUDA path, called through the drop-in Python class -> ctypes -> C ABI,
against (a) the committed fp64 outputs of the unmodified reference and (b) the pinned fp64 oracle
on seeded inputs.  Tolerance (north_star): 1e-5 relative, per term, energies and gradients."""
import numpy as np
import pytest
import torch

from conftest import atominfo_from, rel
from oracle import energy_oracle as eo

pytestmark = pytest.mark.gpu
TOL = 1e-5
TERMS = ("bond", "angle", "torsion", "nb")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def energy_objs(dev, golden_topo):
    from molearn_b200 import TorchProteinEnergy
    cache = {}

    def get(case, NB="repulsive"):
        if (case, NB) not in cache:
            cache[(case, NB)] = TorchProteinEnergy(None, atominfo_from(golden_topo(case)), device=dev, NB=NB)
        return cache[(case, NB)]
    return get


def per_term_grads(tpe, x):
    """[4,B,N,3] gradient of each term's batch sum, through autograd (exercises the weighted backward)."""
    out = []
    for k in range(4):
        xx = x.clone().requires_grad_(True)
        e = tpe.energy_per_conformation(xx)
        (g,) = torch.autograd.grad(e[:, k].sum(), xx)
        out.append(g.permute(0, 2, 1).double().cpu().numpy())
    return np.stack(out)


def test_energies_and_gradients_vs_reference(case, dev, energy_objs, golden_energy):
    en = golden_energy(case)
    tpe = energy_objs(case)
    x = torch.from_numpy(en["x"]).to(dev).permute(0, 2, 1)         # [B,3,N] view of [B,N,3] storage
    e = tpe.energy_per_conformation(x).double().cpu().numpy()
    for k, name in enumerate(TERMS):
        err = np.abs(e[:, k] - en["e64"][:, k]) / np.abs(en["e64"][:, k])
        assert err.max() < TOL, (name, err)
    gf = list(en["grad_frames"])
    g = per_term_grads(tpe, x[gf])
    for k, name in enumerate(TERMS):
        for q in range(len(gf)):
            assert rel(g[k, q], en["g64"][q, k]) < TOL, (name, gf[q], rel(g[k, q], en["g64"][q, k]))


def test_reference_api_conventions(dev, energy_objs, golden_energy):
    """get_energy: bonded batch means + NB batch sum; _roll_...: batch sums (SURVEY E-4)."""
    en = golden_energy("bbcb")
    tpe = energy_objs("bbcb")
    x = torch.from_numpy(en["x"][:8]).to(dev).permute(0, 2, 1).contiguous()   # reference layout [B,3,N]
    s = en["e64"][:8].sum(0)
    b, a, t, nb = (float(v) for v in tpe.get_energy(x, nonbonded=True))
    assert abs(b - s[0] / 8) < TOL * s[0] / 8 and abs(a - s[1] / 8) < TOL * s[1] / 8
    assert abs(t - s[2] / 8) < TOL * s[2] / 8 and abs(nb - s[3]) < TOL * abs(s[3])
    assert len(tpe.get_energy(x)) == 3
    bs, as_, ts = (float(v) for v in tpe._roll_bond_angle_torsion_loss(x))
    assert abs(bs - s[0]) < TOL * s[0] and abs(as_ - s[1]) < TOL * s[1] and abs(ts - s[2]) < TOL * s[2]
    assert abs(float(tpe._nb_loss(x)) - s[3]) < TOL * abs(s[3])
    assert abs(float(tpe._conv_bond_loss(x)) - s[0]) < TOL * s[0]
    # known answers of SURVEY 8c (frames 0-7, fp32 reference run)
    assert abs(b - 812.5279) < 2e-2 and abs(a - 712.2919) < 2e-2 and abs(t - 2845.1143) < 5e-2


def test_total_gradient_one_backward(dev, energy_objs, golden_energy):
    """loss = sum of all four terms: the fused forward gradient, one backward."""
    en = golden_energy("bbcb")
    tpe = energy_objs("bbcb")
    gf = list(en["grad_frames"])
    x = torch.from_numpy(en["x"][gf]).to(dev).permute(0, 2, 1).contiguous().requires_grad_(True)
    b, a, t, nb = tpe.get_energy(x, nonbonded=True)
    (b * len(gf) + a * len(gf) + t * len(gf) + nb).backward()
    want = en["g64"].sum(1)
    got = x.grad.permute(0, 2, 1).double().cpu().numpy()
    for q in range(len(gf)):
        assert rel(got[q], want[q]) < TOL


def test_weighted_backward(dev, energy_objs, golden_energy):
    """Arbitrary upstream weights per conformation and per term."""
    en = golden_energy("bbcb")
    tpe = energy_objs("bbcb")
    gf = list(en["grad_frames"])
    w = torch.tensor([[1.0, -2.0, 0.5, 3.0], [0.0, 1.0, 1.0, 1.0], [2.0, 2.0, 2.0, 2.0], [0.3, 0.0, 0.0, -1.0],
                      [1.0, 1.0, 1.0, 0.0]], device=dev)
    x = torch.from_numpy(en["x"][gf]).to(dev).permute(0, 2, 1).requires_grad_(True)
    e = tpe.energy_per_conformation(x)
    (e * w).sum().backward()
    got = x.grad.permute(0, 2, 1).double().cpu().numpy()
    wn = w.double().cpu().numpy()
    for q in range(len(gf)):
        want = sum(wn[q, k] * en["g64"][q, k] for k in range(4))
        scale = sum(abs(wn[q, k]) * np.linalg.norm(en["g64"][q, k]) for k in range(4))
        assert np.linalg.norm(got[q] - want) < TOL * scale


def test_layouts_agree(dev, energy_objs, golden_energy):
    """[B,3,N] contiguous, permuted [B,N,3] storage, and a decoder-like padded slice."""
    en = golden_energy("bbcb")
    tpe = energy_objs("bbcb")
    xn = torch.from_numpy(en["x"][[0, 9, 12]]).to(dev)              # [B,N,3]
    B, N, _ = xn.shape
    a = xn.permute(0, 2, 1).contiguous()                              # SoA
    b = xn.permute(0, 2, 1)                                           # AoS view
    pad = torch.full((B, N + 87, 3), float("nan"), device=dev)
    pad[:, :N] = xn
    c = pad[:, :N, :].permute(0, 2, 1)                                # sliced decoder output
    outs, grads = [], []
    for xx in (a, b, c):
        xx = xx.detach().requires_grad_(True)
        e = tpe.energy_per_conformation(xx)
        e.sum().backward()
        outs.append(e.detach().cpu().numpy())
        grads.append(xx.grad.permute(0, 2, 1).cpu().numpy())
    for o, g in zip(outs[1:], grads[1:]):
        assert np.allclose(o, outs[0], rtol=2e-6)
        assert rel(g, grads[0]) < 2e-6


def test_clashing_structures_vs_oracle(dev, energy_objs, golden_topo, golden_energy):
    """Garbage-decode regime: heavy noise puts many pairs below 1.9 A and 0.4 A (warp branch)."""
    topo = golden_topo("bbcb")
    tpe = energy_objs("bbcb")
    rng = np.random.default_rng(11)
    base = golden_energy("bbcb")["x"][0]
    xs = np.stack([base + rng.normal(scale=s, size=base.shape).astype(np.float32) for s in (0.3, 2.0, 5.0)])
    xs[2] *= 0.2                                                      # collapsed blob: massive clashes
    x = torch.from_numpy(xs).to(dev).permute(0, 2, 1)
    e = tpe.energy_per_conformation(x).double().cpu().numpy()
    g = per_term_grads(tpe, x)
    for q in range(len(xs)):
        eo_e, eo_g = eo.energy_all(xs[q], topo)
        for k, name in enumerate(TERMS):
            assert abs(e[q, k] - eo_e[k]) < TOL * abs(eo_e[k]), (name, q, e[q, k], eo_e[k])
            assert rel(g[k, q], eo_g[k]) < TOL, (name, q, rel(g[k, q], eo_g[k]))


def test_full_lj_mode_vs_oracle(dev, energy_objs, golden_topo, golden_energy):
    topo = golden_topo("bbcb")
    tpe = energy_objs("bbcb", "full")
    xs = golden_energy("bbcb")["x"][[1, 10]]
    x = torch.from_numpy(xs).to(dev).permute(0, 2, 1).requires_grad_(True)
    e = tpe.energy_per_conformation(x)
    e[:, 3].sum().backward()
    got_g = x.grad.permute(0, 2, 1).double().cpu().numpy()
    ex = eo.exclusions(topo)
    for q in range(2):
        want_e, want_g = eo.nonbonded(xs[q], topo["atom_R"], topo["atom_e"], topo["atom_charges"], ex, full=True)
        assert abs(float(e[q, 3]) - want_e) < TOL * abs(want_e)
        assert rel(got_g[q], want_g) < TOL
    # _cdist_nb_full on a 'repulsive' object and vice versa
    rep = energy_objs("bbcb")
    assert abs(float(rep._cdist_nb_full(x.detach())) - float(e[:, 3].sum())) < 1e-5 * abs(float(e[:, 3].sum()))


def test_edge_batches(dev, energy_objs, golden_energy):
    tpe = energy_objs("bbcb")
    N = tpe.n_atoms
    e = tpe.energy_per_conformation(torch.empty(0, 3, N, device=dev))
    assert e.shape == (0, 4)
    x1 = torch.from_numpy(golden_energy("bbcb")["x"][:1]).to(dev).permute(0, 2, 1)
    e1 = tpe.energy_per_conformation(x1)
    big = x1.expand(37, 3, N)                                         # stride_b = 0 broadcast batch
    eb = tpe.energy_per_conformation(big)
    assert torch.allclose(eb, e1.expand(37, 4), rtol=1e-6)
    with pytest.raises(ValueError):
        tpe.energy_per_conformation(torch.zeros(1, 3, N + 1, device=dev))
    with pytest.raises(RuntimeError):
        tpe.energy_per_conformation(torch.zeros(1, 3, N))
    with pytest.raises(TypeError):
        tpe.energy_per_conformation(torch.zeros(1, 3, N, device=dev, dtype=torch.float64))
    with pytest.raises(TypeError):
        tpe.energy_and_gradient(torch.zeros(1, 3, N, device=dev, dtype=torch.float16))


def test_nonfinite_inputs_stay_in_their_conformation(dev, energy_objs, golden_energy):
    """NaN / inf propagate to that conformation only (the trainer's inf->1e35 / nansum guards rely on it)."""
    tpe = energy_objs("bbcb")
    xs = golden_energy("bbcb")["x"][:3].copy()
    xs[1, 100, 0] = np.nan
    x = torch.from_numpy(xs).to(dev).permute(0, 2, 1).requires_grad_(True)
    e = tpe.energy_per_conformation(x)
    e.nansum().backward()
    e = e.detach().cpu().numpy()
    assert np.isfinite(e[0]).all() and np.isfinite(e[2]).all()
    assert np.isnan(e[1]).all()
    g = x.grad.cpu().numpy()
    assert np.isfinite(g[0]).all() and np.isfinite(g[2]).all()


def test_large_batch_properties(dev, energy_objs, golden_energy):
    """BASELINE batch sizes: translation / rotation invariance and sum of forces = 0 at B=64,
    and a B=4096 batch equals 4096 copies of the small one (checksum of checksums)."""
    en = golden_energy("bbcb")
    tpe = energy_objs("bbcb")
    base = torch.from_numpy(en["x"]).to(dev)                         # [14,N,3]
    g = torch.Generator(device="cpu").manual_seed(0)
    idx = torch.arange(64) % base.shape[0]
    xb = (base[idx] + 0.1 * torch.randn(64, base.shape[1], 3, generator=g).to(dev)).contiguous()
    x = xb.permute(0, 2, 1).requires_grad_(True)
    e = tpe.energy_per_conformation(x)
    e.sum().backward()
    net = x.grad.double().sum(dim=2)                                  # [B,3] net force: zero up to fp32 rounding
    assert float((net.abs() / x.grad.double().abs().sum(dim=2)).max()) < 1e-6
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g).double())
    moved = ((xb.double() - xb.double().mean(1, keepdim=True)) @ q.to(dev)).float().contiguous()
    e2 = tpe.energy_per_conformation(moved.permute(0, 2, 1))
    assert torch.allclose(e, e2, rtol=2e-5)
    big = xb.repeat(64, 1, 1).permute(0, 2, 1)                        # B = 4096
    eb = tpe.energy_per_conformation(big)
    assert torch.equal(eb.view(64, 64, 4)[0], eb.view(64, 64, 4)[63])
    assert torch.allclose(eb.view(64, 64, 4).double().sum(0), 64 * e.double(), rtol=1e-6)


def test_general_torsion_phase_kernel(dev, golden_topo, golden_energy, monkeypatch):
    """ff14SB has only 0 / pi phases, so the product instantiates the cosine-only torsion series; the general
    (sine + cosine) instantiation is forced here and must give the same answers."""
    from molearn_b200 import TorchProteinEnergy
    monkeypatch.setenv("MLP_FORCE_TORSION_SINES", "1")
    tpe = TorchProteinEnergy(None, atominfo_from(golden_topo("bbcb")), device=dev)
    en = golden_energy("bbcb")
    gf = list(en["grad_frames"])
    x = torch.from_numpy(en["x"][gf]).to(dev).permute(0, 2, 1).requires_grad_(True)
    e = tpe.energy_per_conformation(x)
    e[:, 2].sum().backward()
    got = x.grad.permute(0, 2, 1).double().cpu().numpy()
    for q, f in enumerate(gf):
        assert abs(float(e[q, 2]) - en["e64"][f, 2]) < TOL * en["e64"][f, 2]
        assert rel(got[q], en["g64"][q, 2]) < TOL


def test_chunked_evaluation_matches(dev, energy_objs, golden_energy, monkeypatch):
    """Batches whose scratch memory exceeds the cap are evaluated in chunks: same bits as one call."""
    from molearn_b200.loss_functions import torch_protein_energy as tp
    tpe = energy_objs("bbcb")
    x = torch.from_numpy(golden_energy("bbcb")["x"]).to(dev).permute(0, 2, 1)
    e1, g1 = tpe.energy_and_gradient(x)
    monkeypatch.setattr(tp._DeviceEnergy, "MAX_WORKSPACE_BYTES", 3 * 2 ** 20)   # forces chunks of a few conformations
    e2, g2 = tpe.energy_and_gradient(x)
    assert torch.equal(e1, e2) and torch.equal(g1, g2)


def test_tiny_systems_vs_oracle(dev, param_dir):
    """One residue (4 atoms, no torsion), a dipeptide, and a 3-residue peptide with all hydrogens: single tiles,
    inert padding records, mostly-masked non-bonded tile."""
    from molearn_b200 import TorchProteinEnergy
    from oracle import topology_oracle as to
    rng = np.random.default_rng(2)
    at, _, _, _ = to.parse_forcefield(param_dir)
    systems = {
        "gly": [("N", "GLY", 1), ("CA", "GLY", 1), ("C", "GLY", 1), ("O", "GLY", 1)],
        "ala_ser": [(a, "ALA", 1) for a in ("N", "CA", "CB", "C", "O")] + [(a, "SER", 2) for a in ("N", "CA", "CB", "OG", "C", "O")],
        "full_h": [(a, r, k + 1) for k, r in enumerate(("ALA", "TRP", "PRO")) for a in at[r]],
    }
    for name, atoms in systems.items():
        info = np.array(atoms, dtype=object)
        n = len(info)
        topo = to.build_topology(param_dir, list(info[:, 0]), list(info[:, 1]), [int(r) for r in info[:, 2]])
        xs = (np.cumsum(rng.normal(scale=1.2, size=(3, n, 3)), axis=1)).astype(np.float32)
        tpe = TorchProteinEnergy(None, info, device=dev)
        x = torch.from_numpy(xs).to(dev).permute(0, 2, 1)
        e = tpe.energy_per_conformation(x).double().cpu().numpy()
        g = per_term_grads(tpe, x)
        for q in range(3):
            we, wg = eo.energy_all(xs[q], topo)
            for k, term in enumerate(TERMS):
                if abs(we[k]) > 0:
                    assert abs(e[q, k] - we[k]) < TOL * abs(we[k]), (name, term, q, e[q, k], we[k])
                else:
                    assert e[q, k] == 0.0, (name, term)
                if np.linalg.norm(wg[k]) > 0:
                    assert rel(g[k, q], wg[k]) < TOL, (name, term, q, rel(g[k, q], wg[k]))


def test_ten_thousand_atoms_vs_oracle(dev, golden_topo):
    """BASELINE config 4 system (5 chains, 10 715 atoms, 84 x 84 non-bonded tiles): energies and gradient of one
    noisy conformation against the fp64 oracle."""
    from molearn_b200 import TorchProteinEnergy
    from molearn_b200.data import tile_system
    g = load_golden_murd_bbcb()
    tinfo, x = tile_system(g[0], g[1], 5, batch=1, noise=0.3, seed=5)
    tpe = TorchProteinEnergy(None, tinfo, device=dev)
    topo = tpe.topology.as_dict()
    xd = torch.from_numpy(x).to(dev).permute(0, 2, 1).requires_grad_(True)
    e = tpe.energy_per_conformation(xd)
    e.sum().backward()
    want_e, want_g = eo.energy_all(x[0], topo)
    got = e.detach().double().cpu().numpy()[0]
    for k, name in enumerate(TERMS):
        assert abs(got[k] - want_e[k]) < TOL * abs(want_e[k]), (name, got[k], want_e[k])
    assert rel(xd.grad.permute(0, 2, 1).double().cpu().numpy()[0], want_g.sum(0)) < TOL


def load_golden_murd_bbcb():
    from conftest import load_golden
    g = load_golden("murd_input")
    keep = np.isin(g["names"], ("N", "CA", "CB", "C", "O"))
    info = atominfo_from({"in_names": g["names"][keep], "in_resnames": g["resnames"][keep], "in_resids": g["resids"][keep]})
    return info, np.ascontiguousarray(g["coords"][:, keep])


def test_coincident_atoms_follow_reference_semantics(dev, energy_objs, golden_topo, golden_energy):
    """Two bonded atoms on the same point: the bond term is finite (K r0^2), every angle with that arm is 0/0 = NaN,
    torsions around it are atan2(0,0) = 0 -> finite - the same per-term outcome as the reference's forward pass."""
    topo = golden_topo("bbcb")
    tpe = energy_objs("bbcb")
    xs = golden_energy("bbcb")["x"][:2].copy()
    i, j = (int(v) for v in topo["bond_idxs"][700])
    xs[1, j] = xs[1, i]
    e = tpe.energy_per_conformation(torch.from_numpy(xs).to(dev).permute(0, 2, 1)).double().cpu().numpy()
    want, _ = eo.energy_all(xs[1], topo, want_grad=False)
    assert np.isfinite(e[0]).all()
    assert np.isnan(want[1]) and np.isnan(e[1, 1])                       # angle: NaN in both
    for k in (0, 2, 3):                                                  # bond, torsion, nb: finite and equal
        assert np.isfinite(want[k]) and abs(e[1, k] - want[k]) < TOL * abs(want[k]), (k, e[1, k], want[k])


def test_random_peptides_vs_oracle(dev, param_dir):
    """Seeded random peptides (random residue sequences over all 28 templates; backbone+CB, heavy-atom and
    all-atom selections; 150-1400 atoms so that terms, halos and exclusions cross the 256-atom bonded tiles and
    the 128-atom non-bonded tiles at arbitrary offsets), random-walk coordinates: every term and its gradient
    against the fp64 oracle, topology against the Python restatement."""
    from molearn_b200 import TorchProteinEnergy
    from oracle import topology_oracle as to
    at, _, _, _ = to.parse_forcefield(param_dir)
    residues = sorted(at)
    rng = np.random.default_rng(2026)
    selections = {"bbcb": lambda a: a in ("N", "CA", "CB", "C", "O"), "heavy": lambda a: not a.startswith("H"), "all": lambda a: True}
    for trial in range(6):
        sel_name = ("bbcb", "heavy", "all")[trial % 3]
        n_res = int(rng.integers(30, 110))
        seq = [residues[i] for i in rng.integers(0, len(residues), n_res)]
        atoms = [(a, r, k + 1) for k, r in enumerate(seq) for a in at[r] if selections[sel_name](a)]
        info = np.array(atoms, dtype=object)
        n = len(info)
        topo = to.build_topology(param_dir, list(info[:, 0]), list(info[:, 1]), [int(r) for r in info[:, 2]])
        tpe = TorchProteinEnergy(None, info, device=dev)
        assert np.array_equal(tpe.topology.torsion_idxs, topo["torsion_idxs"]) and np.array_equal(tpe.topology.exclusions, to.exclusion_pairs(topo))
        xs = np.cumsum(rng.normal(scale=1.1, size=(2, n, 3)), axis=1).astype(np.float32)
        x = torch.from_numpy(xs).to(dev).permute(0, 2, 1)
        e = tpe.energy_per_conformation(x).double().cpu().numpy()
        g = per_term_grads(tpe, x)
        for q in range(2):
            we, wg = eo.energy_all(xs[q], topo)
            for k, term in enumerate(TERMS):
                assert abs(e[q, k] - we[k]) < TOL * abs(we[k]), (trial, sel_name, n, term, e[q, k], we[k])
                assert rel(g[k, q], wg[k]) < TOL, (trial, sel_name, n, term, rel(g[k, q], wg[k]))


def test_two_streams_one_handle_bit_identical(dev, energy_objs, golden_energy):
    """Independent batches alternating over two CUDA streams through ONE handle (bench.py's two-stream leg): the
    handle's private side stream and events must keep every batch's kernels ordered - results bit-identical to the
    one-stream evaluation."""
    tpe = energy_objs("bbcb")
    x0 = torch.from_numpy(golden_energy("bbcb")["x"]).to(dev)
    gen = torch.Generator(device="cpu").manual_seed(5)
    xs = [(x0.repeat(8, 1, 1) + 0.05 * torch.randn((8 * x0.shape[0],) + tuple(x0.shape[1:]), generator=gen).to(dev)).permute(0, 2, 1)
          for _ in range(6)]
    want = [tpe.energy_and_gradient(x) for x in xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    for rep in range(3):
        got = []
        for i, x in enumerate(xs):
            with torch.cuda.stream(streams[i % 2]):
                got.append(tpe.energy_and_gradient(x))
        torch.cuda.synchronize()
        for (e0, g0), (e1, g1) in zip(want, got):
            assert torch.equal(e0, e1) and torch.equal(g0, g1)


def test_conformation_extent_limit_is_reported(dev, energy_objs, golden_energy):
    """The bonded kernel addresses one conformation with 32-bit element offsets; a layout that spans 2^31 elements
    or more is refused with an error (never evaluated wrongly)."""
    import ctypes
    from molearn_b200 import _lib
    tpe = energy_objs("bbcb")
    dv = tpe._dev()
    x = torch.from_numpy(golden_energy("bbcb")["x"][:1]).to(dev).permute(0, 2, 1).contiguous()
    e = torch.empty((1, 4), device=dev)
    L = _lib.lib()
    nbytes = L.mlp_energy_workspace_bytes(dv._h, 1, _lib.TERM_BONDED, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rc = L.mlp_energy_eval(dv._h, x.data_ptr(), 1, 0, 1, 2 ** 31 // (tpe.n_atoms - 1) + 1, e.data_ptr(), None, 0, 0, 0, None,
                           _lib.TERM_BONDED, 0, ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
    assert rc == _lib.E_ARG and b"32-bit" in L.mlp_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)


@pytest.mark.parametrize("knob", ["MLP_NB_PACKED=0", "MLP_NO_OVERLAP=1", "MLP_BONDED_TILE=64", "MLP_BONDED_TILE=40", "MLP_NB_BLOCKS=8", "MLP_NB_BLOCKS=2",
                                  "MLP_NB_FIRST=1", "MLP_NB_LUT=0", "MLP_NB_WS_MB=1", "MLP_NB_BAND_KB=96"])
def test_alternative_code_paths_agree(knob, dev, golden_topo, golden_energy, monkeypatch):
    """The library's run-time switches (read when a handle is created) select other code paths of the same maths: the
    scalar pair kernel, no side-stream overlap, smaller bonded tiles (other work-item matchings and halos).  Each must
    meet the same tolerance against the reference's fp64 vectors."""
    from molearn_b200 import TorchProteinEnergy, _lib
    name, value = knob.split("=")
    monkeypatch.setenv(name, value)
    tpe = TorchProteinEnergy(None, atominfo_from(golden_topo("bbcb")), device=dev)
    if name == "MLP_BONDED_TILE":
        assert tpe._dev().info()["bonded_tile_atoms"] == int(value)
    # backbone + CB has four Lennard-Jones classes (+ the padding class): the type-pair table is the default path
    assert tpe._dev().info()["nb_table_classes"] == (0 if name == "MLP_NB_LUT" else 5)
    en = golden_energy("bbcb")
    gf = list(en["grad_frames"])
    x = torch.from_numpy(en["x"][gf]).to(dev).permute(0, 2, 1)
    if name == "MLP_NB_BAND_KB":    # 96 KB = 32 tile pairs per band: the 170 tile pairs of MurD in six row bands (the schedule of large systems)
        dv, n0 = tpe._dev(), tpe._dev().launches
        dv.eval(x, torch.empty(len(gf), 4, device=dev), torch.empty_like(x), None, _lib.TERM_ALL)
        assert dv.launches - n0 > 1 + 2 * 4                            # bonded + (pair, reduction) per band
    if name == "MLP_NB_WS_MB":      # 1 MB of partial forces = one conformation per pair launch: a chunked batch
        dv, n0 = tpe._dev(), tpe._dev().launches
        dv.eval(x, torch.empty(len(gf), 4, device=dev), torch.empty_like(x), None, _lib.TERM_ALL)
        per = dv.info()["nb_tile_pairs"] * 2 * 3 * 128 * 4               # bytes of partial forces per conformation
        chunk = max(1, (1 << 20) // per)
        assert chunk < len(gf) and dv.launches - n0 == 1 + 2 * -(-len(gf) // chunk)      # bonded + (pair, reduction) per chunk
    e = tpe.energy_per_conformation(x).double().cpu().numpy()
    for k, term in enumerate(TERMS):
        err = np.abs(e[:, k] - en["e64"][gf][:, k]) / np.abs(en["e64"][gf][:, k])
        assert err.max() < TOL, (knob, term, err)
    g = per_term_grads(tpe, x)
    for k, term in enumerate(TERMS):
        for q in range(len(gf)):
            assert rel(g[k, q], en["g64"][q, k]) < TOL, (knob, term, q)


def test_full_lj_mode_vs_reference_vectors(dev, energy_objs):
    """NB='full' against fp64 runs of the reference's own ``_cdist_nb_full`` (torch_protein_energy.py:224-230;
    tests/golden/energy_full_*.npz, written by oracle/make_golden.py --full-nb)."""
    from conftest import load_golden
    for case in ("bbcb", "pep28"):
        en = load_golden("energy_full_" + case)
        tpe = energy_objs(case, "full")
        gf = list(en["grad_frames"])
        x = torch.from_numpy(en["x"]).to(dev).permute(0, 2, 1).requires_grad_(True)
        e = tpe.energy_per_conformation(x)
        e[:, 3].sum().backward()
        got_e = e[:, 3].detach().double().cpu().numpy()
        assert np.all(np.abs(got_e - en["e64"]) < TOL * np.abs(en["e64"])), (case, got_e, en["e64"])
        got_g = x.grad.permute(0, 2, 1).double().cpu().numpy()
        for q, f in enumerate(gf):
            assert rel(got_g[f], en["g64"][q]) < TOL, (case, f, rel(got_g[f], en["g64"][q]))
        # the reference's batch-sum call on the same object
        assert abs(float(tpe._nb_loss(x.detach())) - en["e64"].sum()) < TOL * abs(en["e64"].sum())


def test_zero_upstream_weight_stops_nonfinite_terms(dev, energy_objs, golden_topo, golden_energy):
    """A term or a conformation the loss masked out (zero upstream gradient) must not reach dL/dx as NaN * 0:
    (a) a conformation with a NaN coordinate and weight 0 gets an exactly zero gradient, the others their exact one;
    (b) a conformation whose ANGLE term is NaN (coincident bonded atoms, finite coordinates) but whose loss only
        uses bond + nb gets the finite bond + nb gradient (the trainer's nansum over the three bonded sums does this,
        torch_physics_trainer.py:40-42)."""
    topo = golden_topo("bbcb")
    tpe = energy_objs("bbcb")
    xs = golden_energy("bbcb")["x"][:3].copy()
    xs[1, 100, 0] = np.nan
    i, j = (int(v) for v in topo["bond_idxs"][700])
    xs[2, j] = xs[2, i]
    x = torch.from_numpy(xs).to(dev).permute(0, 2, 1).requires_grad_(True)
    e = tpe.energy_per_conformation(x)
    w = torch.tensor([[1.0, 1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 1.0]], device=dev)
    torch.nansum(e * w).backward()
    g = x.grad.permute(0, 2, 1).double().cpu().numpy()
    assert np.isfinite(g).all()
    assert (g[1] == 0).all()
    _, g0 = eo.energy_all(xs[0], topo)
    assert rel(g[0], g0.sum(0)) < TOL
    # the oracle's bond gradient carries its own 0/0 on the coincident pair (forward-mode duals of |v| at v = 0);
    # autograd - and the kernel - give 0 there, so compare away from those two atoms
    _, g2 = eo.energy_all(xs[2], topo)
    want = g2[0] + g2[3]
    keep = np.ones(len(want), bool)
    keep[[i, j]] = False
    assert np.isfinite(want[keep]).all() and rel(g[2][keep], want[keep]) < TOL
    # the same through bonded-only energies and three separately weighted sums (the trainer's call)
    x2 = torch.from_numpy(xs).to(dev).permute(0, 2, 1).requires_grad_(True)
    b, a, t = tpe._roll_bond_angle_torsion_loss(x2)
    assert torch.isnan(a)
    torch.nansum(torch.stack([b, a, t])).backward()
    assert torch.isfinite(x2.grad[[0, 2]]).all()


# ---- BASELINE config 5: the 51 432-atom system (24 chains) ---------------------------------------------------------
@pytest.fixture(scope="module")
def system_51k(dev):
    from molearn_b200 import TorchProteinEnergy
    from molearn_b200.data import tile_system
    g = load_golden_murd_bbcb()
    tinfo, x = tile_system(g[0], g[1], 24, batch=1, noise=0.2, seed=9)
    tpe = TorchProteinEnergy(None, tinfo, device=dev)
    assert tpe.n_atoms == 51432
    return tpe, tpe.topology.as_dict(), x


def _check_51k(tpe, topo, xs, dev, tols, label):
    xd = torch.from_numpy(xs).to(dev).permute(0, 2, 1).requires_grad_(True)
    e = tpe.energy_per_conformation(xd)
    e.nansum().backward()
    got = e.detach().double().cpu().numpy()[0]
    want_e, want_g = eo.energy_all(xs[0], topo)
    for k, name in enumerate(TERMS):
        if np.isnan(want_e[k]):
            assert np.isnan(got[k]), (label, name, got[k])
        else:
            assert abs(got[k] - want_e[k]) <= tols[k] * abs(want_e[k]), (label, name, got[k], want_e[k], abs(got[k] - want_e[k]) / abs(want_e[k]))
    return got, want_e, xd.grad.permute(0, 2, 1).double().cpu().numpy()[0], want_g


def test_51k_atoms_md_like_vs_oracle(dev, system_51k):
    """MD-like conformation of the config-5 system (0.2 A noise): every term and the total gradient <= 1e-5."""
    tpe, topo, x = system_51k
    got, want, g, wg = _check_51k(tpe, topo, x, dev, [TOL] * 4, "md-like")
    assert rel(g, wg.sum(0)) < TOL


def test_51k_atoms_collapsed_vs_oracle(dev, system_51k):
    """A collapsed structure (the regime of an untrained decoder: everything within a few Angstrom, E_nb ~ 1e10, the
    pair kernel's exact-only loop) WITHOUT coincident atoms: every term and the gradient <= 1e-5."""
    tpe, topo, x = system_51k
    rng = np.random.default_rng(51)
    xs = ((x - x.mean(axis=1, keepdims=True)) * 0.15 + rng.normal(scale=0.25, size=x.shape)).astype(np.float32)
    got, want, g, wg = _check_51k(tpe, topo, xs, dev, [TOL] * 4, "collapsed")
    assert want[3] > 1e7
    assert rel(g, wg.sum(0)) < TOL


def test_51k_atoms_coincident_atoms_semantics(dev, system_51k):
    """The same collapsed structure with three pairs of bonded atoms put on the same point (an untrained decoder emits
    such duplicates): bond finite, angle NaN (0/0), torsion finite with phi = atan2(0, 0) = 0, non-bonded finite - the
    per-term outcome of the reference's forward pass, here and in the oracle; the finite terms still agree."""
    tpe, topo, x = system_51k
    rng = np.random.default_rng(52)
    xs = ((x - x.mean(axis=1, keepdims=True)) * 0.15 + rng.normal(scale=0.25, size=x.shape)).astype(np.float32)
    for k in (100, 20000, 51000):
        i, j = (int(v) for v in topo["bond_idxs"][k])
        xs[0, j] = xs[0, i]
    got, want, _, _ = _check_51k(tpe, topo, xs, dev, [TOL, TOL, 1e-4, TOL], "coincident")
    assert np.isfinite(got[0]) and np.isnan(got[1]) and np.isfinite(got[2]) and np.isfinite(got[3])


@pytest.mark.parametrize("schedule", ["whole list", "block items"])
def test_51k_atoms_other_pair_schedules(dev, system_51k, monkeypatch, schedule):
    """The config-5 system runs the tile-pair kernel band by band through its tile rows (four launches of <= 64 MB of
    partial forces per conformation).  Two other schedules must give the same numbers: the whole tile-pair list in one
    launch (MLP_NB_BANDS=0 MLP_NB_BLOCKS=0: 81 405 tile pairs are more than one grid dimension holds, the pair index
    spans blockIdx.y and blockIdx.z) and the block work items (MLP_NB_BANDS=0).  Each against the oracle at 1e-5 and
    against the default at 1e-6."""
    from molearn_b200 import TorchProteinEnergy, _lib
    from molearn_b200.data import tile_system
    tpe_def, topo, x = system_51k
    L = _lib.lib()
    per_conf = lambda t: L.mlp_energy_workspace_bytes(t._dev()._h, 1, _lib.TERM_ALL, 1)
    assert per_conf(tpe_def) < 80 << 20                                   # banded: <= 64 MB of partial forces
    monkeypatch.setenv("MLP_NB_BANDS", "0")
    if schedule == "whole list":
        monkeypatch.setenv("MLP_NB_BLOCKS", "0")
    g = load_golden_murd_bbcb()
    tinfo, _ = tile_system(g[0], g[1], 24, batch=1, noise=0.2, seed=9)
    tpe = TorchProteinEnergy(None, tinfo, device=dev)
    assert tpe._dev().info()["nb_tile_pairs"] > 65535
    assert (per_conf(tpe) > 200 << 20) == (schedule == "whole list")
    got, want, g1, wg = _check_51k(tpe, topo, x, dev, [TOL] * 4, "md-like, " + schedule)
    assert rel(g1, wg.sum(0)) < TOL
    got2, _, g2, _ = _check_51k(tpe_def, topo, x, dev, [TOL] * 4, "md-like, row bands")
    assert abs(got[3] - got2[3]) <= 1e-6 * abs(got2[3]) and rel(g1, g2) < 1e-6


def test_block_item_schedule_matches_tile_pair_schedule(dev, monkeypatch):
    """The two schedules of the pair term - one warp per tile pair (nb_kernel2) and one warp per block of tile pairs
    (nb_block_kernel: rows kept in registers, column forces accumulated in shared memory, cp.async double buffering) -
    on the 10 715-atom system: MD-like, noisy and collapsed conformations, both NB modes; against each other and
    against the fp64 oracle.  The block schedule needs less scratch."""
    from molearn_b200 import TorchProteinEnergy, _lib
    from molearn_b200.data import tile_system
    g = load_golden_murd_bbcb()
    tinfo, x = tile_system(g[0], g[1], 5, batch=3, noise=0.3, seed=6)
    x[2] = (x[2] - x[2].mean(0)) * 0.2 + np.random.default_rng(4).normal(scale=0.2, size=x[2].shape)      # collapsed
    xd = torch.from_numpy(x.astype(np.float32)).to(dev).permute(0, 2, 1)
    for NB in ("repulsive", "full"):
        res, ws = {}, {}
        for mode in ("0", "4", "8"):
            monkeypatch.setenv("MLP_NB_BLOCKS", mode)
            tpe = TorchProteinEnergy(None, tinfo, device=dev, NB=NB)
            res[mode] = tpe.energy_and_gradient(xd)
            ws[mode] = _lib.lib().mlp_energy_workspace_bytes(tpe._dev()._h, 1, _lib.TERM_ALL, 1)
            topo = tpe.topology.as_dict()
        (e0, g0), (e1, g1) = res["0"], res["8"]
        for other in ("4", "8"):
            assert torch.allclose(e0, res[other][0], rtol=2e-6), (NB, other, e0, res[other][0])
            assert float((g0 - res[other][1]).norm() / g0.norm()) < 2e-6
        assert ws["8"] < 0.3 * ws["0"] and ws["4"] < 0.4 * ws["0"], ws
        for qi in (0, 2):
            we, wg = eo.energy_all(x[qi].astype(np.float32), topo, full=NB == "full")
            got = e1[qi].double().cpu().numpy()
            assert abs(got[3] - we[3]) < TOL * abs(we[3]), (NB, qi, got[3], we[3])
            assert rel(g1[qi].permute(1, 0).double().cpu().numpy(), wg.sum(0)) < TOL
    # energy-only and weighted-backward paths of the block schedule
    monkeypatch.setenv("MLP_NB_BLOCKS", "2")
    tpe = TorchProteinEnergy(None, tinfo, device=dev)
    xr = xd.detach().clone().requires_grad_(True)
    e = tpe.energy_per_conformation(xr)
    (e * torch.tensor([[1.0, 0.5, 2.0, -1.0], [0.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 3.0]], device=dev)).sum().backward()
    with torch.no_grad():
        e_only = tpe.energy_per_conformation(xd)
    assert torch.equal(e.detach(), e_only)
    assert (xr.grad[1] == 0).all() and torch.isfinite(xr.grad).all()


def test_graphed_step_equals_eager(dev, energy_objs, golden_energy):
    """GraphedEnergyStep: H2D + energy + backward + D2H captured into one CUDA graph per pinned buffer (through the
    autograd Function), replayed over two streams - bit-identical to the eager call, also after the buffers change."""
    from molearn_b200.loss_functions import GraphedEnergyStep
    tpe = energy_objs("bbcb")
    x0 = torch.from_numpy(golden_energy("bbcb")["x"])                  # [14,N,3]
    bufs = [x0.clone().pin_memory(), (x0 + 0.05).pin_memory(), x0.flip(0).contiguous().pin_memory()]
    step = GraphedEnergyStep(tpe, bufs)
    for rep in range(2):
        for k in range(3):
            step.submit(k)
        for k in range(3):
            e, g = step.result(k)
            xe = bufs[k].to(dev).permute(0, 2, 1).requires_grad_(True)
            ee = tpe.energy_per_conformation(xe)
            ee.sum().backward()
            assert torch.equal(e, ee.detach().cpu()) and torch.equal(g, xe.grad.permute(0, 2, 1).cpu()), (rep, k)
        bufs[1].add_(0.01)                                              # the loader writes new batches into the same pinned buffers
        bufs[2].mul_(1.001)
    # dE/dx left on the device (a consumer on the GPU), three streams with two priority levels: the same numbers
    step2 = GraphedEnergyStep(tpe, bufs, n_streams=3, gradient_to_host=False, priority_levels=2)
    for k in range(3):
        step2.submit(k)
    for k in range(3):
        e, g = step2.result(k)
        assert g.is_cuda and not e.is_cuda
        step.submit(k)
        e1, g1 = step.result(k)
        step2.synchronize()
        assert torch.equal(e, e1) and torch.equal(g.cpu(), g1), k
    with pytest.raises(ValueError):
        GraphedEnergyStep(tpe, [x0])                                    # not pinned
