"""Ad-hoc large-N check: BASELINE config 5 system (51 432 atoms), one MD-like conformation and one decoded
(random-init FoldingNet) structure, kernels vs the fp64 oracle, term by term."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import murd
from molearn_b200 import TorchProteinEnergy
from molearn_b200.data import tile_system
from molearn_b200.models import AutoEncoder
from oracle import energy_oracle as eo
dev = torch.device("cuda:0")
info, frames = murd()
tinfo, x = tile_system(info, frames, 24, batch=1, noise=0.2, seed=9)
tpe = TorchProteinEnergy(None, tinfo, device=dev)
topo = tpe.topology.as_dict()
print("N", tpe.n_atoms, tpe._dev().info())
def check(xs, label):
    xd = torch.from_numpy(xs).to(dev).permute(0, 2, 1).requires_grad_(True)
    e = tpe.energy_per_conformation(xd); e.nansum().backward()
    got = e.detach().double().cpu().numpy()[0]
    t = time.time(); we, wg = eo.energy_all(xs[0], topo); print(label, "oracle s", round(time.time() - t, 1))
    for k, n in enumerate(("bond", "angle", "torsion", "nb")):
        print("  ", n, got[k], we[k], "rel", abs(got[k] - we[k]) / abs(we[k]) if np.isfinite(we[k]) and we[k] != 0 else "n/a")
    g = xd.grad.permute(0, 2, 1).double().cpu().numpy()[0]
    if np.isfinite(wg).all():
        print("   grad rel", np.linalg.norm(g - wg.sum(0)) / np.linalg.norm(wg.sum(0)))
check(x, "md-like")
torch.manual_seed(0)
net = AutoEncoder(out_points=x.shape[1]).to(dev).eval()
xs = torch.from_numpy(x).to(dev)
std, mean = float(xs.std()), float(xs.mean())
with torch.no_grad():
    z = torch.tensor([[-1.1, -1.1], [0.3, -0.7], [1.1, 1.1], [0.0, 0.0]], device=dev)
    dec = (net.decode(z)[:, :x.shape[1]] * std + mean)
    e = tpe.energy_per_conformation(dec.permute(0, 2, 1))
print("decoded energies", e.cpu().numpy())
bad = (~torch.isfinite(e).all(dim=1)).nonzero().flatten().tolist()
pick = bad[0] if bad else 0
check(dec[pick:pick + 1].contiguous().cpu().numpy(), f"decoded point {pick} ({'non-finite' if bad else 'finite'})")


def torsion_energy_torch(xyz, dtype):
    """The reference's torsion formula (torch_protein_energy.py:295-304) on explicit lists, in the given dtype:
    how far plain fp32 arithmetic sits from fp64 on this structure (conditioning of collapsed geometries)."""
    xt = torch.from_numpy(xyz).to(dev, dtype)
    idx = torch.from_numpy(topo["torsion_idxs"]).to(dev).long()
    para = torch.from_numpy(topo["torsion_para"]).to(dev).float().to(dtype)      # fp32-rounded parameters, like the reference
    b1 = xt[idx[:, 0]] - xt[idx[:, 1]]; b2 = xt[idx[:, 1]] - xt[idx[:, 2]]; b3 = xt[idx[:, 3]] - xt[idx[:, 2]]
    c32 = torch.linalg.cross(b3, b2); c12 = torch.linalg.cross(b1, b2)
    phi = torch.atan2((b2 * torch.linalg.cross(c32, c12)).sum(-1), b2.norm(dim=-1) * (c12 * c32).sum(-1))
    e = (para[:, 1] / para[:, 0]) * (1 + torch.cos(para[:, 3] * phi[:, None] - para[:, 2]))
    return float(e.sum().double())


xd = dec[pick].contiguous().cpu().numpy()
e64, e32 = torsion_energy_torch(xd, torch.float64), torsion_energy_torch(xd, torch.float32)
print("torsion energy of that structure with the reference formula: fp64", e64, "fp32", e32, "rel", abs(e32 - e64) / abs(e64))
