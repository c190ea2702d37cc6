"""Step time of the full energy + gradient call (MurD backbone+CB) over batch sizes; device-resident inputs."""
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
base = make_batch(frames, 0).to(dev)
print(f"{'batch':>6s} {'ms/step':>9s} {'conf/s':>10s} {'bonded-only ms':>15s}")
for B in (1, 4, 8, 16, 32, 64, 128, 256, 1024, 4096):
    x = base.repeat((B + 63) // 64, 1, 1)[:B].contiguous()
    x = x + 0.01 * torch.randn_like(x)
    g = torch.empty_like(x); e = torch.empty((B, 4), device=dev)
    out = []
    for mask in (_lib.TERM_ALL, _lib.TERM_BONDED):
        n = max(20, min(400, 40000 // B))
        for _ in range(5):
            dv.eval(x.permute(0, 2, 1), e, g.permute(0, 2, 1), None, mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            dv.eval(x.permute(0, 2, 1), e, g.permute(0, 2, 1), None, mask)
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / n)
    print(f"{B:6d} {out[0]:9.4f} {B / out[0] * 1e3:10.0f} {out[1]:15.4f}")
