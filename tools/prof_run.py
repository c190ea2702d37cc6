"""Small driver for ncu captures: a few hot-path evaluations, nothing else.
usage: python tools/prof_run.py [bonded|nb|all] [B] [case]   (case: bbcb | tile5 | tileK)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch, murd  # noqa: E402
from molearn_b200 import TorchProteinEnergy, _lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
case = sys.argv[3] if len(sys.argv) > 3 else "bbcb"
dev = torch.device("cuda:0")
info, frames = murd()
if case.startswith("tile"):
    from molearn_b200.data import tile_system
    tinfo, xs = tile_system(info, frames, int(case[4:]), batch=B, noise=0.05, seed=1234)
    tpe = TorchProteinEnergy(None, tinfo, device=dev)
    x = torch.from_numpy(xs).to(dev)
else:
    tpe = TorchProteinEnergy(None, info, device=dev)
    x = make_batch(frames, 0, min(B, 64)).to(dev)
    if B > 64:
        x = x.repeat(B // 64, 1, 1)
dv = tpe._dev()
g = torch.empty_like(x)
e = torch.empty((x.shape[0], 4), device=dev)
mask = {"bonded": _lib.TERM_BONDED, "nb": _lib.TERM_NB, "all": _lib.TERM_ALL}[what]
for _ in range(3):
    dv.eval(x.permute(0, 2, 1), e, g.permute(0, 2, 1), None, mask)
torch.cuda.synchronize()
print("ok", e[0].tolist())
