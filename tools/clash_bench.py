"""Pair-kernel time in the clash regime (collapsed structures: most pairs closer than 1.9 A) vs the normal regime."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch, murd  # noqa: E402
from molearn_b200 import TorchProteinEnergy, _lib  # noqa: E402

dev = torch.device("cuda:0")
info, frames = murd()
tpe = TorchProteinEnergy(None, info, device=dev)
dv = tpe._dev()
B = 256
base = make_batch(frames, 0).to(dev).repeat(B // 64, 1, 1)
gen = torch.Generator(device="cpu").manual_seed(3)
cases = {"md-like": base,
         "1 A noise": base + torch.randn(base.shape, generator=gen).to(dev),
         "collapsed blob (sigma 2 A)": 2.0 * torch.randn(base.shape, generator=gen).to(dev)}
for name, x in cases.items():
    x = x.contiguous()
    g = torch.empty_like(x); e = torch.empty((B, 4), device=dev)
    for grad in (False, True):
        gg = g.permute(0, 2, 1) if grad else None
        for _ in range(3):
            dv.eval(x.permute(0, 2, 1), e, gg, None, _lib.TERM_NB)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dv.eval(x.permute(0, 2, 1), e, gg, None, _lib.TERM_NB)
        e1.record(); torch.cuda.synchronize()
        print(f"{name:28s} grad={int(grad)}  {e0.elapsed_time(e1) / 10:8.3f} ms per batch of {B}   E_nb[0] = {float(e[0, 3]):.6g}")
