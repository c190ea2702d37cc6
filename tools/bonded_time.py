"""Time bonded_kernel<grad> of whatever library MOLEARN_B200_LIB points at: MurD backbone+CB, B conformations
(default 4096, the roofline_bonded configuration), CUDA events inside the C ABI.
usage: python tools/bonded_time.py [tag] [B]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch, murd  # noqa: E402
from molearn_b200 import TorchProteinEnergy, _lib  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else ""
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
info, frames = murd()
tpe = TorchProteinEnergy(None, info, device=dev)
x = make_batch(frames, 0, 64).to(dev).repeat(B // 64, 1, 1)
x = x + 0.01 * torch.randn_like(x)
dv = tpe._dev()
g = torch.empty_like(x)
e = torch.empty((B, 4), device=dev)
out = {"tag": tag, "B": B}
for grad in (True, False):
    gg = g.permute(0, 2, 1) if grad else None
    for _ in range(3):
        dv.eval(x.permute(0, 2, 1), e, gg, None, _lib.TERM_BONDED)
    torch.cuda.synchronize()
    dv.profile(True)
    for _ in range(20):
        dv.eval(x.permute(0, 2, 1), e, gg, None, _lib.TERM_BONDED)
    pr = dv.profile_read()
    dv.profile(False)
    ms = pr["bonded"][0] / max(pr["bonded"][1], 1)
    key = "grad" if grad else "energy"
    out[key + "_ms"] = round(ms, 5)
    if grad:
        out["hbm_frac"] = round(24.0 * x.shape[1] * B / (ms * 1e-3) / 1e9 / 6539.5, 4)
out["esum"] = float(e.double().sum())
out["gsum"] = float(g.double().abs().sum())
print(json.dumps(out))
