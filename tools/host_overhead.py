"""Host-side cost per call of the public entry points (wall time of the Python call, GPU work queued asynchronously).
usage: python tools/host_overhead.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch, murd  # noqa: E402
from molearn_b200 import TorchProteinEnergy, _lib  # noqa: E402

dev = torch.device("cuda:0")
info, frames = murd()
tpe = TorchProteinEnergy(None, info, device=dev)
dv = tpe._dev()
x = make_batch(frames, 0).to(dev)
xp = x.permute(0, 2, 1)
g = torch.empty_like(x); e = torch.empty((x.shape[0], 4), device=dev)

def raw():
    dv.eval(xp, e, g.permute(0, 2, 1), None, _lib.TERM_ALL)
def direct():
    tpe.energy_and_gradient(xp)
def autograd():
    xx = xp.detach().requires_grad_(True)
    en = tpe.energy_per_conformation(xx)
    en.sum().backward()
    return xx.grad
def bonded_only_bs8():
    dv.eval(xp[:8], e[:8], g[:8].permute(0, 2, 1), None, _lib.TERM_BONDED)

for name, fn, n in (("raw dv.eval", raw, 300), ("energy_and_gradient", direct, 300), ("autograd", autograd, 300), ("bonded B=8", bonded_only_bs8, 2000)):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    # (a) host cost: queue depth kept small by syncing every 8 calls, time only the calls
    t_host = 0.0
    for i in range(n):
        t0 = time.perf_counter(); fn(); t_host += time.perf_counter() - t0
        if i % 4 == 3: torch.cuda.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n): fn()
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"{name:22s} host {1e6 * t_host / n:7.1f} us/call   back-to-back {1e6 * t_all / n:7.1f} us/call")
