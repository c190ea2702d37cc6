"""CPU: the fp64 C restatement (oracle/energy_oracle.c) against fp64 runs of the unmodified
reference (tests/golden/energy_*.npz).  This is what pins the checker used by the GPU tests."""
import numpy as np

import pytest
import torch

from conftest import load_golden, rel
from oracle import energy_oracle as eo


def test_energy_oracle_matches_reference(case, golden_topo, golden_energy):
    topo, en = golden_topo(case), golden_energy(case)
    gf = list(en["grad_frames"])
    frames = sorted(set(gf) | {0, len(en["x"]) - 1})
    for i in frames:
        e, g = eo.energy_all(en["x"][i], topo, want_grad=i in gf)
        # bonded: same formulas, same fp32-rounded parameters -> fp64 round-off only
        assert np.all(np.abs(e[:3] - en["e64"][i, :3]) <= 1e-12 * np.abs(en["e64"][i, :3]))
        # NB: the reference's A_ij matrix is built in fp32 (cdist noise ~1e-6 on the LJ part)
        assert abs(e[3] - en["e64"][i, 3]) <= 2e-7 * abs(en["e64"][i, 3])
        if i in gf:
            gg = en["g64"][gf.index(i)]
            for k in range(3):
                assert rel(g[k], gg[k]) < 1e-12
            assert rel(g[3], gg[3]) < 2e-6


def test_reference_fp32_is_within_its_own_noise(golden_energy):
    """Documents the distance between the reference's fp32 result and its fp64 result."""
    en = golden_energy("bbcb")
    s64 = en["e64"].sum(0)
    assert np.all(np.abs(en["ref32_batch_sums"] - s64) <= 1e-4 * np.abs(s64))


def test_rmsd_oracle():
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(50, 3)), rng.normal(size=(50, 3))
    assert abs(eo.rmsd(a, b) - np.sqrt(((a - b) ** 2).sum(1).mean())) < 1e-14


@pytest.mark.parametrize("case", ["bbcb", "pep28"])
def test_energy_oracle_full_nb_matches_reference(case, golden_topo):
    """The oracle's NB='full' branch (attractive -B/w^6 added) against fp64 runs of the reference's
    ``_cdist_nb_full`` (torch_protein_energy.py:224-230), vectors from oracle/make_golden.py --full-nb."""
    topo, en = golden_topo(case), load_golden("energy_full_" + case)
    gf = list(en["grad_frames"])
    for i in range(len(en["x"])):
        e, g = eo.nonbonded(en["x"][i], topo["atom_R"], topo["atom_e"], topo["atom_charges"], eo.exclusions(topo),
                            full=True, want_grad=i in gf)
        assert abs(e - en["e64"][i]) <= 2e-7 * abs(en["e64"][i]), (i, e, en["e64"][i])
        if i in gf:
            assert rel(g, en["g64"][gf.index(i)]) < 2e-6


def _port(topo, dtype, **kw):
    from oracle.torch_port import TorchCPUPort
    return TorchCPUPort(topo, dtype=dtype, **kw)


def test_torch_port_is_pinned_to_the_reference(golden_topo, golden_energy):
    """oracle/torch_port.py is the timed CPU baseline (``--impl reference`` and every ``cpu_baseline``): its fp32
    batch sums and total gradient reproduce the reference's own fp32 run, its fp64 run the reference's fp64 energies."""
    topo, en = golden_topo("bbcb"), golden_energy("bbcb")
    x = torch.from_numpy(en["x"]).permute(0, 2, 1).contiguous()
    e32, g32 = _port(topo, torch.float32).energy_and_grad(x)
    assert np.all(np.abs(e32.double().numpy() - en["ref32_batch_sums"]) <= 2e-6 * np.abs(en["ref32_batch_sums"]))
    assert rel(g32.permute(0, 2, 1).numpy(), en["ref32_grad_total"]) < 1e-6
    sel = [0, 9, 12]
    e64, _ = _port(topo, torch.float64).energy_and_grad(x[sel])
    want = en["e64"][sel].sum(0)
    assert np.all(np.abs(e64.numpy() - want) <= 1e-10 * np.abs(want))


def test_torch_port_full_nb(golden_topo):
    topo, en = golden_topo("pep28"), load_golden("energy_full_pep28")
    x = torch.from_numpy(en["x"]).permute(0, 2, 1).contiguous()
    port = _port(topo, torch.float64, nb_full=True)
    for i in range(len(x)):
        assert abs(float(port.nonbonded(x[i:i + 1].double())) - en["e64"][i]) <= 1e-9 * abs(en["e64"][i])
