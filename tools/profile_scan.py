"""Kernel-time breakdown of a small latent-grid scan (torch profiler, CUDA activities)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import murd  # noqa: E402
from molearn_b200.analysis import MolearnAnalysis  # noqa: E402
from molearn_b200.data import CoordinateSet  # noqa: E402
from molearn_b200.models import AutoEncoder  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

dev = torch.device("cuda:0")
info, frames = murd()
data = CoordinateSet(info, frames).prepare_dataset()
torch.manual_seed(0)
net = AutoEncoder(out_points=frames.shape[1]).to(dev)
ma = MolearnAnalysis(batch_size=256)
ma.set_network(net)
ma.set_dataset("murd", data)
tg = torch.from_numpy(frames)
ma.setup_grid(samples=32, padding=0.1)
ma.scan_scores(targets=tg)
torch.cuda.synchronize()
ma.setup_grid(samples=64, padding=0.1)
t0 = time.perf_counter(); ma.scan_scores(targets=tg); torch.cuda.synchronize(); t1 = time.perf_counter()
print("64x64 scan:", (t1 - t0) * 1e3, "ms =", (t1 - t0) * 1e3 / 16, "ms per batch of 256")
ma.setup_grid(samples=32, padding=0.1)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ma.scan_scores(targets=tg)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=80))
