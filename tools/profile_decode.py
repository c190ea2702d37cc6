"""Kernel-time breakdown of one FoldingNet decode batch (torch profiler, CUDA activities)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molearn_b200.models import AutoEncoder  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = AutoEncoder(out_points=2145).to(dev).eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
z = torch.rand(B, 2, device=dev) * 2 - 1
with torch.no_grad():
    for _ in range(3):
        net.decode(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        net.decode(z)
    e1.record(); torch.cuda.synchronize()
    print("decode ms per batch", e0.elapsed_time(e1) / 5, "per structure", e0.elapsed_time(e1) / 5 / B)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net.decode(z)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=90))
from molearn_b200.models import FusedDecoder  # noqa: E402
fused = FusedDecoder(net.decoder)
with torch.no_grad():
    for _ in range(3):
        fused(z)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fused(z)
    e1.record(); torch.cuda.synchronize()
    print("FUSED decode ms per batch", e0.elapsed_time(e1) / 5, "per structure", e0.elapsed_time(e1) / 5 / B, "fused conv+relu:", fused.fused_ok)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fused(z)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=90))
