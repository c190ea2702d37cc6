"""TEST INFRASTRUCTURE ONLY - regenerate ``tests/golden/*.npz`` from the reference itself.

Run in the build container (needs ``/root/reference``):  ``python oracle/make_golden.py``.
The reference is loaded unmodified through ``oracle/ref_loader.py`` (shims listed there); the
vectors written here are what pins the oracle restatements (``oracle/topology_oracle.py``,
``oracle/energy_oracle.c``) and, through them, the CUDA path, on machines that have no
reference tree.  Everything is seeded; inputs are stored next to outputs.

Files
  murd_input.npz            the reference's own test fixture test/MurD_test.pdb reduced to
                            arrays: names/resnames/resids [3286], coords [16,3286,3] f32
  topo_<case>.npz           index lists, fp64 parameter arrays, per-atom NB parameters and a
                            banded + random sample of the dense NB matrices
  energy_<case>.npz         inputs x [B,N,3] f32 and, per term (bond, angle, torsion, nb):
                            fp32 reference batch values, fp64 per-conformation energies and
                            fp64 gradients
  energy_full_<case>.npz    NB='full' (_cdist_nb_full): fp64 energies / gradients, the 1-4 pair list and the 1-4
                            scaled parameter arrays (bbcb, pep28; ``--full-nb`` writes only these)
Cases: bbcb (N,CA,CB,C,O -> 2145 atoms), heavy (all 3286), tile2 (two bbcb tiles as two
chains, SURVEY 8d config 4 construction, 4288 -> 4286 atoms), pep28 (one residue of each of the 28
amino12.lib templates with all hydrogens, 444 atoms, random-walk coordinates).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader, topology_oracle  # noqa: E402
from molearn_b200.data.pdb_reader import read_pdb_frames, select_atoms  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BB = ["N", "CA", "CB", "C", "O"]


def info_arrays(info):
    return dict(names=np.array([str(a) for a in info[:, 0]]),
                resnames=np.array([str(a) for a in info[:, 1]]),
                resids=np.array([int(a) for a in info[:, 2]], dtype=np.int64))


def make_tiles(info, coords, n_tiles, frame_of_tile):
    """tile = bbcb MurD minus the last residue's C and O, resids
    consecutive across tiles, tile t translated on an 80 A lattice."""
    last = info[-1, 2]
    keep = np.array([not (r == last and a in ("C", "O")) for a, _, r in info])
    tinfo, tcoords = info[keep], coords[:, keep]
    nres = int(tinfo[-1, 2]) - int(tinfo[0, 2]) + 1
    infos, xs = [], []
    for t in range(n_tiles):
        ti = tinfo.copy()
        ti[:, 2] = [int(r) + t * nres for r in tinfo[:, 2]]
        infos.append(ti)
        shift = 80.0 * np.array([t % 3, (t // 3) % 3, t // 9], dtype=np.float32)
        xs.append(tcoords[frame_of_tile(t)] + shift)
    return np.concatenate(infos), np.concatenate(xs, axis=0)


def topology_case(tag, info, frame0):
    frame = torch.from_numpy(frame0.T.copy()).float()  # [3,N]
    obj, cap, info_m = ref_loader.build_reference(frame, info, pair_exclusions=True)
    n = len(info)
    d = dict(
        **{"in_" + k: v for k, v in info_arrays(info).items()},
        **{"out_" + k: v for k, v in info_arrays(info_m).items()},
        amber_types=np.array([str(a[0]) for a in cap["atom_names"]]),
        bond_idxs=np.asarray(cap["bond_idxs"], dtype=np.int64),
        angle_idxs=np.asarray(cap["angle_idxs"], dtype=np.int64),
        torsion_idxs=np.asarray(cap["torsion_idxs"], dtype=np.int64),
        bond_14_idxs=np.asarray(cap["bond_14_idxs"], dtype=np.int64),
        bond_para=np.asarray(cap["bond_para"], dtype=np.float64),
        angle_para=np.asarray(cap["angle_para"], dtype=np.float64),
        torsion_para=np.asarray(cap["torsion_para"], dtype=np.float64),
        atom_charges=np.asarray(cap["atom_charges"].double().numpy()
                                if torch.is_tensor(cap["atom_charges"]) else cap["atom_charges"],
                                dtype=np.float64),
        atom_R=cap["atom_R"].numpy().astype(np.float32),
        atom_e=cap["atom_e"].numpy().astype(np.float32),
    )
    A, Q = obj.vdw_A.numpy(), obj.q1q2.numpy()
    d["nb_nnz_A"] = np.int64((A != 0).sum())
    d["nb_nnz_q"] = np.int64((Q != 0).sum())
    d["nb_sum_A"] = np.float64(A.astype(np.float64).sum())
    d["nb_sum_q"] = np.float64(Q.astype(np.float64).sum())
    # full band |j-i| <= 20 (covers every exclusion) + a seeded random sample of the rest
    ii, jj = [], []
    for k in range(1, 21):
        i = np.arange(0, n - k)
        ii.append(i)
        jj.append(i + k)
    rng = np.random.default_rng(7)
    ri = rng.integers(0, n, 40000)
    rj = rng.integers(0, n, 40000)
    ii.append(ri)
    jj.append(rj)
    ii, jj = np.concatenate(ii), np.concatenate(jj)
    d["nb_i"], d["nb_j"] = ii.astype(np.int32), jj.astype(np.int32)
    d["nb_A"], d["nb_q"] = A[ii, jj].astype(np.float32), Q[ii, jj].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, f"topo_{tag}.npz"), **d)
    print(tag, "N", n, "bonds", len(d["bond_idxs"]), "angles", len(d["angle_idxs"]),
          "torsions", len(d["torsion_idxs"]), "P", d["torsion_para"].shape[2],
          "nnzA", int(d["nb_nnz_A"]))
    return obj


def energy_case(tag, obj, x_bn3, grad_frames):
    """x_bn3: [B,N,3] float32 numpy."""
    x32 = torch.from_numpy(x_bn3).permute(0, 2, 1).contiguous()  # [B,3,N] the reference layout
    B = x32.shape[0]
    out = dict(x=x_bn3.astype(np.float32))
    # fp32 reference, whole batch, as shipped (roll + matmul-cdist)
    xr = x32.clone().requires_grad_(True)
    b, a, t = obj._roll_bond_angle_torsion_loss(xr)
    nb = obj._nb_loss(xr)
    (b + a + t + nb).backward()
    out["ref32_batch_sums"] = np.array([b.item(), a.item(), t.item(), nb.item()], dtype=np.float64)
    out["ref32_grad_total"] = xr.grad.permute(0, 2, 1).contiguous().numpy()
    be, ae, te = obj.get_energy(x32)
    out["ref32_get_energy"] = np.array([be.item(), ae.item(), te.item()], dtype=np.float64)
    # fp64 reference, one conformation at a time, per term
    obj64 = ref_loader.to_double(obj)
    e64 = np.zeros((B, 4), dtype=np.float64)
    g64 = np.zeros((len(grad_frames), 4, x_bn3.shape[1], 3), dtype=np.float64)
    for i in range(B):
        xi = x32[i:i + 1].double().requires_grad_(True)
        terms = list(obj64._roll_bond_angle_torsion_loss(xi)) + [obj64._nb_loss(xi)]
        for k, e in enumerate(terms):
            e64[i, k] = e.item()
            if i in grad_frames:
                (g,) = torch.autograd.grad(e, xi, retain_graph=(k < 3))
                g64[grad_frames.index(i), k] = g[0].T.numpy()
    out["e64"] = e64
    out["g64"] = g64
    out["grad_frames"] = np.array(grad_frames, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, f"energy_{tag}.npz"), **out)
    print(tag, "fp32 batch sums", out["ref32_batch_sums"], "fp64 sums", e64.sum(0))


def full_nb_case(tag, info, x_bn3, grad_frames):
    """``NB='full'`` (``_cdist_nb_full``, torch_protein_energy.py:224-230: adds the attractive -B/w^6 term) and the 1-4
    scaled parameter arrays the reference computes but never uses (utils.py:833-836): fp64 per-conformation energies and
    gradients from the unmodified reference -> energy_full_<tag>.npz (pins the oracle's ``full`` branch)."""
    frame = torch.from_numpy(x_bn3[0].T.copy()).float()
    obj, cap, _ = ref_loader.build_reference(frame, info, pair_exclusions=True, NB="full")
    x32 = torch.from_numpy(x_bn3).permute(0, 2, 1).contiguous()
    B = x32.shape[0]
    out = dict(x=x_bn3.astype(np.float32))
    xr = x32.clone().requires_grad_(True)
    nb = obj._nb_loss(xr)
    nb.backward()
    out["ref32_batch_sum"] = np.float64(nb.item())
    obj64 = ref_loader.to_double(obj)
    e64 = np.zeros(B)
    g64 = np.zeros((len(grad_frames), x_bn3.shape[1], 3))
    for i in range(B):
        xi = x32[i:i + 1].double().requires_grad_(True)
        e = obj64._nb_loss(xi)
        e64[i] = e.item()
        if i in grad_frames:
            (g,) = torch.autograd.grad(e, xi)
            g64[grad_frames.index(i)] = g[0].T.numpy()
    out["e64"], out["g64"], out["grad_frames"] = e64, g64, np.array(grad_frames, dtype=np.int64)
    as64 = lambda v: np.asarray(v.double().numpy() if torch.is_tensor(v) else v, dtype=np.float64)
    out["bond_14_idxs"] = np.asarray(cap["bond_14_idxs"], dtype=np.int64)
    out["vdw_14R"], out["vdw_14e"], out["q1q2_14"] = as64(cap["vdw_14R"]), as64(cap["vdw_14e"]), as64(cap["q1q2_14"])
    np.savez_compressed(os.path.join(OUT, f"energy_full_{tag}.npz"), **out)
    print(tag, "NB=full fp32 batch sum", out["ref32_batch_sum"], "fp64 sum", e64.sum(), "n14", len(out["bond_14_idxs"]))


def main_full():
    """Round 2 additions; the files written by main() are left untouched."""
    info, X = read_pdb_frames(os.path.join(ref_loader.REF, "test", "MurD_test.pdb"))
    g = torch.Generator().manual_seed(20261020)
    binfo, bX = select_atoms(info, X, BB)
    noise = torch.randn(2, bX.shape[1], 3, generator=g).numpy().astype(np.float32)
    xb = np.concatenate([bX[:3], bX[3:4] + 1.0 * noise[:1], bX[4:5] + 0.1 * noise[1:]])
    full_nb_case("bbcb", binfo, xb, grad_frames=[0, 3, 4])
    at, _, _, _ = topology_oracle.parse_forcefield(ref_loader.PARAM_DIR)
    pinfo = np.array([(name, res, k + 1) for k, res in enumerate(at) for name in at[res]], dtype=object)
    rng = np.random.default_rng(28)
    walk = np.cumsum(rng.normal(scale=0.9, size=(len(pinfo), 3)), axis=0).astype(np.float32)
    xp = np.stack([walk, walk * 1.5, walk * 0.6]).astype(np.float32)
    full_nb_case("pep28", pinfo, xp, grad_frames=[0, 1, 2])


def main():
    os.makedirs(OUT, exist_ok=True)
    info, X = read_pdb_frames(os.path.join(ref_loader.REF, "test", "MurD_test.pdb"))
    np.savez_compressed(os.path.join(OUT, "murd_input.npz"), coords=X, **info_arrays(info))
    g = torch.Generator().manual_seed(20261019)

    # --- bbcb -----------------------------------------------------------------------------
    binfo, bX = select_atoms(info, X, BB)
    obj = topology_case("bbcb", binfo, bX[0])
    noise = torch.randn(4, bX.shape[1], 3, generator=g).numpy().astype(np.float32)
    xb = np.concatenate([bX[:8], bX[:4] + 1.0 * noise, bX[8:10] + 0.1 * noise[:2]])
    obj, _, _ = ref_loader.build_reference(torch.from_numpy(bX[0].T.copy()), binfo)
    energy_case("bbcb", obj, xb, grad_frames=[0, 1, 8, 9, 12])

    # --- all heavy atoms ------------------------------------------------------------------
    obj = topology_case("heavy", info, X[0])
    # NB: B must not be 3 - the reference's dim-less torch.cross (torch_protein_energy.py:
    # 299-301) then picks the batch axis (first axis of size 3) and returns garbage.
    noise = torch.randn(2, X.shape[1], 3, generator=g).numpy().astype(np.float32)
    xh = np.concatenate([X[:2], X[2:4] + 0.5 * noise])
    obj, _, _ = ref_loader.build_reference(torch.from_numpy(X[0].T.copy()), info)
    energy_case("heavy", obj, xh, grad_frames=[0, 2])

    # --- every residue template of amino12.lib, all atoms incl. hydrogens ------------------
    # 28 residues (ALA ... VAL incl. ASH/CYM/CYX/GLH/HID/HIE/HIP/HYP/LYN), template atom order, consecutive
    # resids; coordinates are a seeded random walk (0.9 A steps): not a structure, but finite, and it
    # exercises every atom type / bond / angle / torsion parameter the templates can produce, rings,
    # hydrogens, and the close-contact branch of the non-bonded term.
    at, _, _, _ = topology_oracle.parse_forcefield(ref_loader.PARAM_DIR)
    pinfo = np.array([(name, res, k + 1) for k, res in enumerate(at) for name in at[res]], dtype=object)
    rng = np.random.default_rng(28)
    walk = np.cumsum(rng.normal(scale=0.9, size=(len(pinfo), 3)), axis=0).astype(np.float32)
    xp = np.stack([walk, walk * 1.5, walk + rng.normal(scale=0.3, size=walk.shape).astype(np.float32),
                   walk * 0.6]).astype(np.float32)
    obj = topology_case("pep28", pinfo, xp[0])
    obj, _, _ = ref_loader.build_reference(torch.from_numpy(xp[0].T.copy()), pinfo)
    energy_case("pep28", obj, xp, grad_frames=[0, 1, 3])

    # --- two tiles = two chains -----------------------------------------------------------
    tinfo, tx = make_tiles(binfo, bX, 2, lambda t: t % 16)
    obj = topology_case("tile2", tinfo, tx)
    obj, _, _ = ref_loader.build_reference(torch.from_numpy(tx.T.copy()), tinfo)
    energy_case("tile2", obj, tx[None], grad_frames=[0])


if __name__ == "__main__":
    if "--full-nb" in sys.argv:
        main_full()
    else:
        main()
        main_full()
