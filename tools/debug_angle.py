import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from oracle import energy_oracle as eo
from molearn_b200.data import CoordinateSet
from molearn_b200.analysis import MolearnAnalysis
from molearn_b200.models import AutoEncoder
d = load_golden("murd_input"); keep = np.isin(d["names"], ("N","CA","CB","C","O"))
info = np.empty((int(keep.sum()), 3), dtype=object)
info[:, 0], info[:, 1], info[:, 2] = list(d["names"][keep]), list(d["resnames"][keep]), [int(r) for r in d["resids"][keep]]
data = CoordinateSet(info, np.ascontiguousarray(d["coords"][:, keep])).prepare_dataset()
torch.manual_seed(0)
net = AutoEncoder(out_points=data.dataset.shape[1]).to("cuda:0")
ma = MolearnAnalysis(batch_size=16); ma.set_network(net); ma.set_dataset("train", data); ma.setup_grid(samples=6, padding=0.1)
with torch.no_grad():
    z = ma._encoded["grid"].reshape(-1, 2).to(ma.device)
    dec = (ma.network.decode(z)[:, :ma.n_atoms, :] * ma.stdval + ma.meanval)
x = dec[0:1].contiguous()
tpe = ma.physics()
e = tpe.energy_per_conformation(x.permute(0, 2, 1))
print("kernel", e.cpu().numpy())
topo = load_golden("topo_bbcb")
xn = x[0].double().cpu().numpy()
print("oracle", eo.energy_all(xn, topo, want_grad=False)[0])
print("coords range", xn.min(), xn.max(), "std", xn.std())
A = topo["angle_idxs"]; ap = topo["angle_para"].astype(np.float32).astype(np.float64)
v1 = xn[A[:, 0]] - xn[A[:, 1]]; v2 = xn[A[:, 2]] - xn[A[:, 1]]
n1 = (v1 * v1).sum(1); n2 = (v2 * v2).sum(1)
c = (v1 * v2).sum(1) / np.sqrt(n1 * n2)
th = np.arccos(np.clip(c, -0.999999, 0.999999))
E = ap[:, 1] * (th - ap[:, 0]) ** 2
print("numpy64 angle sum", E.sum())
# fp32 emulation of the kernel
x32 = x[0].cpu().numpy()
w1 = x32[A[:, 0]] - x32[A[:, 1]]; w2 = x32[A[:, 2]] - x32[A[:, 1]]
m1 = (w1 * w1).sum(1, dtype=np.float32); m2 = (w2 * w2).sum(1, dtype=np.float32)
c32 = ((w1 * w2).sum(1, dtype=np.float32) / np.sqrt(m1 * m2)).astype(np.float32)
th32 = np.arccos(np.clip(c32, np.float32(-0.999999), np.float32(0.999999))).astype(np.float32)
E32 = ap[:, 1].astype(np.float32) * (th32 - ap[:, 0].astype(np.float32)) ** 2
print("numpy32 angle sum", E32.sum(dtype=np.float64))
dE = E32.astype(np.float64) - E
o = np.argsort(-np.abs(dE))[:10]
for k in o: print(k, A[k], "c64", c[k], "c32", c32[k], "n1", n1[k], "n2", n2[k], "E64", E[k], "E32", E32[k])
print("min n", n1.min(), n2.min(), "n clamp", (np.abs(c) > 0.999999).sum())
