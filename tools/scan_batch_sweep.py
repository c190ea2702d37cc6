"""Scan throughput (MurD, decode + energy + RMSD to 16 frames) over the decode batch size."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import murd  # noqa: E402
from molearn_b200.analysis import MolearnAnalysis  # noqa: E402
from molearn_b200.data import CoordinateSet  # noqa: E402
from molearn_b200.models import AutoEncoder  # noqa: E402

dev = torch.device("cuda:0")
info, frames = murd()
data = CoordinateSet(info, frames).prepare_dataset()
torch.manual_seed(0)
net = AutoEncoder(out_points=frames.shape[1]).to(dev)
tg = torch.from_numpy(frames)
for bs in (128, 256, 512, 1024):
    ma = MolearnAnalysis(batch_size=bs)
    ma.set_network(net)
    ma.set_dataset("murd", data)
    ma.setup_grid(samples=64, padding=0.1)
    ma.scan_scores(targets=tg)
    torch.cuda.synchronize()
    ma.setup_grid(samples=128, padding=0.1)
    t0 = time.perf_counter(); ma.scan_scores(targets=tg); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"batch {bs:5d}: {16384 / (t1 - t0):9.0f} points/s   peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
