"""cProfile of the host side of the autograd call path (energy_per_conformation + sum + backward)."""
import cProfile, os, pstats, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch, murd  # noqa: E402
from molearn_b200 import TorchProteinEnergy  # noqa: E402

dev = torch.device("cuda:0")
info, frames = murd()
tpe = TorchProteinEnergy(None, info, device=dev)
x = make_batch(frames, 0).to(dev).permute(0, 2, 1)

def one():
    xx = x.detach().requires_grad_(True)
    e = tpe.energy_per_conformation(xx)
    e.sum().backward()
    return xx.grad

for _ in range(50): one()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(400):
    one()
    if i % 4 == 3: torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
