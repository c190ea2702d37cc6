"""CPU: the fp64 C restatement (oracle/energy_oracle.c) against fp64 runs of the unmodified
reference (tests/golden/energy_*.npz).  This is what pins the checker used by the GPU tests."""
import numpy as np

from conftest import rel
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
