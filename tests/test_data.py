"""CPU: the minimal PDB / DCD readers and the synthetic multi-chain builder that feed tests and benches."""
import struct

import numpy as np

from conftest import atominfo_from, load_golden
from molearn_b200.data import CoordinateSet, read_dcd, read_pdb_frames, select_atoms, tile_system


def write_pdb(path, info, frames):
    with open(path, "w") as fh:
        for f in frames:
            for k, ((name, res, rid), xyz) in enumerate(zip(info, f)):
                fh.write("ATOM  %5d %-4s %3s A%4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % (k + 1, name, res, rid, *xyz))
            fh.write("END\n")


def write_dcd(path, frames):
    """Little-endian CHARMM DCD, no unit cell: the layout of the reference's test/MurD_test.dcd."""
    nset, natom = frames.shape[:2]
    icntrl = [0] * 20
    icntrl[0], icntrl[19] = nset, 24
    hdr = b"CORD" + struct.pack("<20i", *icntrl)
    title = struct.pack("<i", 1) + b"x".ljust(80)
    with open(path, "wb") as fh:
        for rec in (hdr, title, struct.pack("<i", natom)):
            fh.write(struct.pack("<i", len(rec)) + rec + struct.pack("<i", len(rec)))
        for f in frames:
            for c in range(3):
                body = f[:, c].astype("<f4").tobytes()
                fh.write(struct.pack("<i", len(body)) + body + struct.pack("<i", len(body)))


def test_pdb_and_dcd_readers_round_trip(tmp_path):
    g = load_golden("murd_input")
    info = atominfo_from({"in_names": g["names"][:60], "in_resnames": g["resnames"][:60], "in_resids": g["resids"][:60]})
    frames = g["coords"][:3, :60]
    write_pdb(tmp_path / "t.pdb", info, frames)
    info2, x2 = read_pdb_frames(str(tmp_path / "t.pdb"))
    assert np.array_equal(info2, info) and x2.shape == (3, 60, 3)
    assert np.allclose(x2, frames, atol=6e-4)                          # %8.3f
    write_dcd(tmp_path / "t.dcd", frames)
    assert np.array_equal(read_dcd(str(tmp_path / "t.dcd")), frames)   # DCD stores fp32: exact
    sel_info, sel_x = select_atoms(info2, x2, ["N", "CA", "C"])
    assert set(sel_info[:, 0]) == {"N", "CA", "C"} and sel_x.shape[1] == len(sel_info)
    cs = CoordinateSet(sel_info, sel_x).prepare_dataset()
    assert abs(float(cs.dataset.mean())) < 1e-5 and abs(float(cs.dataset.std()) - 1) < 1e-5
    assert np.allclose(cs.dataset.numpy() * cs.std + cs.mean, sel_x, atol=1e-4)


def test_tile_system_matches_config4_counts():
    """5 tiles of MurD backbone+CB minus the last residue's C, O: the N and term counts BASELINE.md states."""
    from molearn_b200 import Topology
    g = load_golden("murd_input")
    keep = np.isin(g["names"], ("N", "CA", "CB", "C", "O"))
    info = atominfo_from({"in_names": g["names"][keep], "in_resnames": g["resnames"][keep], "in_resids": g["resids"][keep]})
    tinfo, x = tile_system(info, g["coords"][:, keep], 5, batch=2, noise=0.05)
    assert x.shape == (2, 10715, 3) and len(tinfo) == 10715
    t = Topology(tinfo)
    assert (len(t.bond_idxs), len(t.angle_idxs), len(t.torsion_idxs)) == (10710, 14870, 16845)
    # no bond between tiles: every bond stays inside one 2143-atom tile
    assert np.all(t.bond_idxs[:, 0] // 2143 == t.bond_idxs[:, 1] // 2143)
