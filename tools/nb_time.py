"""Time the hot-path kernels of whatever library MOLEARN_B200_LIB points at (CUDA events inside the C ABI).
usage: python tools/nb_time.py [tag]   -> one JSON line: pair kernel / bonded / reduce ms for MurD B=64 and the
10 715-atom system (B=32), plus the whole-call time of the headline batch."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch, murd  # noqa: E402
from molearn_b200 import TorchProteinEnergy, _lib  # noqa: E402
from molearn_b200.data import tile_system  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else ""
dev = torch.device("cuda:0")
info, frames = murd()
out = {"tag": tag}


def run(tpe, x, reps, key):
    dv = tpe._dev()
    g = torch.empty_like(x)
    e = torch.empty((x.shape[0], 4), device=dev)
    for _ in range(3):
        dv.eval(x.permute(0, 2, 1), e, g.permute(0, 2, 1), None, _lib.TERM_ALL)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dv.eval(x.permute(0, 2, 1), e, g.permute(0, 2, 1), None, _lib.TERM_ALL)
    e1.record()
    torch.cuda.synchronize()
    out[key + "_call_ms"] = e0.elapsed_time(e1) / reps
    dv.profile(True)
    for _ in range(reps):
        dv.eval(x.permute(0, 2, 1), e, g.permute(0, 2, 1), None, _lib.TERM_ALL)
    pr = dv.profile_read()
    dv.profile(False)
    for k in ("nb", "bonded", "nb_reduce"):
        out[f"{key}_{k}_ms"] = pr[k][0] / max(pr[k][1], 1)
    pairs = tpe.topology.n_nb_pairs * x.shape[0]
    out[key + "_frac"] = 31 * pairs / (out[key + "_nb_ms"] * 1e-3) / 1e12 / 74.45
    out[key + "_esum"] = float(e.double().sum())


tpe = TorchProteinEnergy(None, info, device=dev)
cases = os.environ.get("NB_TIME_CASES", "murd64,n10k32" if os.environ.get("NB_TIME_BIG", "1") == "1" else "murd64").split(",")
for case in cases:
    if case.startswith("murd"):
        Bc = int(case[4:])
        xb = make_batch(frames, 0, min(Bc, 64)).to(dev)
        if Bc > 64:
            xb = xb.repeat(Bc // 64, 1, 1) + 0.01 * torch.randn(Bc, xb.shape[1], 3, device=dev)
        run(tpe, xb, 50 if Bc <= 256 else 5, case)
    elif case.startswith("n10k"):
        Bc = int(case[4:])
        tinfo, x = tile_system(info, frames, 5, batch=Bc, noise=0.05, seed=1234)
        tpe2 = TorchProteinEnergy(None, tinfo, device=dev)
        run(tpe2, torch.from_numpy(x).to(dev), 5, case)
        out[case + "_ws_mb"] = _lib.lib().mlp_energy_workspace_bytes(tpe2._dev()._h, Bc, _lib.TERM_ALL, 1) / 2 ** 20
        del tpe2
    elif case.startswith("n51k"):
        Bc = int(case[4:])
        tinfo, x = tile_system(info, frames, 24, batch=Bc, noise=0.05, seed=1234)
        tpe2 = TorchProteinEnergy(None, tinfo, device=dev)
        run(tpe2, torch.from_numpy(x).to(dev), 3, case)
        out[case + "_ws_mb"] = _lib.lib().mlp_energy_workspace_bytes(tpe2._dev()._h, Bc, _lib.TERM_ALL, 1) / 2 ** 20
        del tpe2
print(json.dumps({k: (round(v, 5) if isinstance(v, float) else v) for k, v in out.items()}))
