"""Largest relative error per energy term and of the non-bonded gradient against the reference's fp64 vectors
(tests/golden) for the four golden systems: how far inside the 1e-5 tolerance of the GPU tests the kernels sit.
usage (GPU box): python tools/parity_margins.py"""
import sys, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden, atominfo_from
from molearn_b200 import TorchProteinEnergy
dev = torch.device('cuda:0')
for case in ('bbcb', 'heavy', 'pep28', 'tile2'):
    topo = load_golden('topo_' + case); en = load_golden('energy_' + case)
    tpe = TorchProteinEnergy(None, atominfo_from(topo), device=dev)
    gf = list(en['grad_frames'])
    x = torch.from_numpy(en['x']).to(dev).permute(0, 2, 1)
    e = tpe.energy_per_conformation(x).double().cpu().numpy()
    err = np.abs(e - en['e64']) / np.abs(en['e64'])
    xg = torch.from_numpy(en['x'][gf]).to(dev).permute(0, 2, 1).requires_grad_(True)
    ee = tpe.energy_per_conformation(xg); ee[:, 3].sum().backward()
    g = xg.grad.permute(0, 2, 1).double().cpu().numpy()
    gerr = [np.linalg.norm(g[q] - en['g64'][q, 3]) / np.linalg.norm(en['g64'][q, 3]) for q in range(len(gf))]
    print(case, 'table classes', tpe._dev().info()['nb_table_classes'], 'max rel err per term', err.max(0), 'nb grad', max(gerr))
