"""Where the time of a config-5 scan batch goes (51 432 atoms, batch 8): decoder vs pair kernel vs the rest."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import murd  # noqa: E402
from molearn_b200.analysis import MolearnAnalysis  # noqa: E402
from molearn_b200.data import CoordinateSet, tile_system  # noqa: E402
from molearn_b200.models import AutoEncoder  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

dev = torch.device("cuda:0")
info, frames = murd()
tinfo, x = tile_system(info, frames, 24, batch=1)
data = CoordinateSet(tinfo, x).prepare_dataset()
torch.manual_seed(0)
net = AutoEncoder(out_points=x.shape[1]).to(dev)
ma = MolearnAnalysis(batch_size=8)
ma.set_network(net)
ma.set_dataset("complex", data)
ma.setup_grid(samples=4, bounds=(-1, 1, -1, 1), padding=0.1)
ma.scan_scores()
torch.cuda.synchronize()
ma.setup_grid(samples=8, bounds=(-1, 1, -1, 1), padding=0.1)
t0 = time.perf_counter(); ma.scan_scores(); torch.cuda.synchronize(); t1 = time.perf_counter()
print("8x8 scan:", (t1 - t0) * 1e3, "ms =", (t1 - t0) * 1e3 / 8, "ms per batch of 8 =", 64 / (t1 - t0), "points/s")
ma.setup_grid(samples=4, bounds=(-1, 1, -1, 1), padding=0.1)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ma.scan_scores()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=80))
