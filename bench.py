#!/usr/bin/env python
"""Benchmark of the molearn physics hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): conformations/s, energy + gradient of all four TorchProteinEnergy terms
(bond, angle, torsion, soft non-bonded), MurD backbone+CB (N = 2145), batch 64, FP32.
One "step" = one batch of 64 conformations through the hot path.

  value     device-resident inputs, C-ABI call (mlp_energy_eval: energies + dE/dx), CUDA events
  e2e       the public Python API (TorchProteinEnergy.energy_per_conformation + backward) with
            pinned HOST buffers: H2D of the batch, D2H of energies and gradient inside the timing
  roofline  the dominant kernel (non-bonded pair kernel) against the FP32 FMA peak, plus the
            bonded kernel against the measured HBM bandwidth at a large batch (roofline_bonded)
  cpu_baseline / --impl reference
            the PyTorch CPU port of the reference algorithm (oracle/torch_port.py, pinned to the reference's own
            vectors by tests/test_oracle_energy.py) on the host cores, on a bounded sample (8 of the 64 conformations)
  torch_cuda_baseline
            the same op chain (what the reference runs with device='cuda') on the same B200: conformations/s,
            kernel launches per step and peak memory, next to the three launches of this library

N > 1 (torchrun): the loss path is "replicas only"  - every rank runs its own batch
of 64, value = total conformations / max-over-ranks time ("weak").  The latent-grid scans, which
do shard, are reported in the "scan" (BASELINE config 3) and "config5_50k_scan" objects and, as short
flat keys (scan_pts_s, scan_ms, scan_checksum, config5_pts_s, ...), at the END of the JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "conformations/sec energy+grad (MurD, batch 64)"
UNIT = "conformations/s"
B = 64
ATOMS = ("N", "CA", "CB", "C", "O")
FLOPS_PER_PAIR = 31          # FMA = 2, MUFU not counted, fast path
CPU_SAMPLE = 8


# ---------------------------------------------------------------------------------------------
def murd():
    d = np.load(os.path.join(ROOT, "tests", "golden", "murd_input.npz"))
    keep = np.isin(d["names"], ATOMS)
    info = np.empty((int(keep.sum()), 3), dtype=object)
    info[:, 0] = [str(s) for s in d["names"][keep]]
    info[:, 1] = [str(s) for s in d["resnames"][keep]]
    info[:, 2] = [int(r) for r in d["resids"][keep]]
    return info, np.ascontiguousarray(d["coords"][:, keep])          # [16,N,3] Angstrom


def make_batch(frames, seed, batch=B):
    """SURVEY 8d config 2 input: the 16 test frames tiled + N(0, 0.1 A) noise, seeded."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.arange(batch) % frames.shape[0]
    x = torch.from_numpy(frames)[idx] + 0.1 * torch.randn(batch, frames.shape[1], 3, generator=g)
    return x.float().contiguous()                                      # [B,N,3]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for l in self.lines:
            p = [s.strip() for s in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ---------------------------------------------------------------------------------------------
def cpu_port_rate(topo_dict, x_bn3, reps, warm=1):
    """conformations/s of the PyTorch CPU port (fwd + bwd, all four terms), all host threads."""
    from oracle.torch_port import TorchCPUPort
    torch.set_num_threads(os.cpu_count() or 1)
    port = TorchCPUPort(topo_dict)
    x = x_bn3.permute(0, 2, 1).contiguous()
    for _ in range(warm):
        port.energy_and_grad(x)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        port.energy_and_grad(x)
        ts.append(time.perf_counter() - t0)
    return x.shape[0] / float(np.median(ts)), float(np.median(ts)), torch.get_num_threads()


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU path (PyTorch port) on all host threads.  Each step is a
    bounded sample of the workload - as many of the batch's 64 conformations (at most 8) as keep the whole
    --steps/--warmup run within a couple of minutes."""
    if rank != 0:
        return
    info, frames = murd()
    # topology arrays as the reference itself produced them (tests/golden/topo_bbcb.npz, written by
    # oracle/make_golden.py from the unmodified reference): this arm loads nothing of the product
    topo = dict(np.load(os.path.join(ROOT, "tests", "golden", "topo_bbcb.npz")))
    assert len(topo["atom_R"]) == len(info)
    from oracle.torch_port import TorchCPUPort
    torch.set_num_threads(os.cpu_count() or 1)
    port = TorchCPUPort(topo)
    full = make_batch(frames, 0)[:CPU_SAMPLE].permute(0, 2, 1).contiguous()
    port.energy_and_grad(full)                                         # first-touch
    t0 = time.perf_counter()
    port.energy_and_grad(full)
    per_conf = (time.perf_counter() - t0) / CPU_SAMPLE
    budget = 120.0 / max(1, args.steps + args.warmup)
    sample = int(max(1, min(CPU_SAMPLE, budget // per_conf)))
    xx = full[:sample].contiguous()
    for _ in range(args.warmup):
        port.energy_and_grad(xx)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.energy_and_grad(xx)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} of the {B} conformations per step (fwd+bwd, 4 terms), PyTorch CPU port of the reference (roll + cdist)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(len(info)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_atoms):
    return {"workload": f"MurD backbone+CB (N={n_atoms}) TorchProteinEnergy bond+angle+torsion+soft-NB energy and dE/dx, batch {B}",
            "batch": B, "n_atoms": n_atoms, "terms": "bond,angle,torsion,nb(repulsive)",
            "l2": "inputs/outputs rotate through a pool of batches larger than the 126 MB L2",
            "parallelism": "replicas (one batch per GPU, no data-path collective)"}


# ---------------------------------------------------------------------------------------------
def rows_checksum(rows):
    """fp64 sum of log1p|v| over every entry of the gathered score rows: equal rows give equal checksums, and the
    collapsed decodes' 1e10-sized energies do not drown the rest (a plain sum would)."""
    r = rows.double()
    return float(torch.log1p(r.abs())[torch.isfinite(r)].sum())       # non-finite entries (0/0 angles of coincident atoms) are counted apart


def rows_nonfinite(rows):
    return int((~torch.isfinite(rows)).sum())


def recheck(ma, rows, n, **kw):
    """Largest relative difference between the gathered rows of a sharded scan and the same grid points scored again
    on THIS rank alone (n points spread over the whole grid, i.e. over every rank's block): the proof, on hardware,
    that N ranks give the surfaces one rank gives."""
    total = rows.shape[0]
    idx = torch.linspace(0, total - 1, n).long().unique()
    again = ma.scan_scores_at(idx, **kw).double()
    ref = rows[idx].double()
    same_nan = torch.isfinite(again) == torch.isfinite(ref)
    d = ((again - ref).abs() / ref.abs().clamp_min(1e-30))[torch.isfinite(ref) & torch.isfinite(again)]
    worst = float(d.max()) if d.numel() else 0.0
    return (worst if bool(same_nan.all()) else float("inf")), int(len(idx))


def torch_cuda_rate(topo_dict, x_bn3, dev, reps, nonbonded=True):
    """The reference's op chain (oracle/torch_port.py: roll groups over [B,3,N] + dense cdist NB) on the SAME GPU -
    what `TorchProteinEnergy(device='cuda')` of the reference runs.  fwd + bwd; CUDA events; kernel launches per step
    from the torch profiler; peak memory from the caching allocator."""
    from oracle.torch_port import TorchCPUPort
    port = TorchCPUPort(topo_dict, device=dev)
    x = x_bn3.permute(0, 2, 1).contiguous().to(dev)
    for _ in range(2):
        port.energy_and_grad(x, nonbonded=nonbonded)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        port.energy_and_grad(x, nonbonded=nonbonded)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    peak = torch.cuda.max_memory_allocated(dev)
    launches = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            port.energy_and_grad(x, nonbonded=nonbonded)
            torch.cuda.synchronize()
        launches = int(sum(ev.count for ev in prof.key_averages() if getattr(ev, "device_type", None) is not None
                           and "cuda" in str(ev.device_type).lower()))
    except Exception:
        pass
    del port
    torch.cuda.empty_cache()
    return {"value": x.shape[0] / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": int(x.shape[0]),
            "kernel_launches_per_step": launches, "peak_memory_bytes": int(peak), "kind": "port",
            "what": "PyTorch port of the reference op chain (" + ("roll groups + dense cdist NB" if nonbonded else "roll groups, bonded terms only")
                    + ") on this GPU, fwd+bwd, fp32"}


def run_scan(dev, rank, world, info, frames, G, barrier):
    """MurD latent-grid scan GxG: decode (FoldingNet, cuDNN) + 4 energy terms + RMSD to the 16 dataset
    frames per point, through MolearnAnalysis.scan_scores; grid points sharded over the ranks and
    gathered with one all_gather.  points/s = G^2 / max-over-ranks time (strong scaling)."""
    import torch.distributed as dist
    from molearn_b200.analysis import MolearnAnalysis
    from molearn_b200.data import CoordinateSet
    from molearn_b200.models import AutoEncoder
    data = CoordinateSet(info, frames).prepare_dataset()
    torch.manual_seed(0)
    # (cuDNN's default heuristics; cudnn.benchmark tries workspaces of > 140 GB on this decoder for an 8 % gain)
    net = AutoEncoder(out_points=frames.shape[1]).to(dev)
    ma = MolearnAnalysis(batch_size=256)
    ma.set_network(net)
    ma.set_dataset("murd", data)
    targets = torch.from_numpy(frames)
    warm = int(np.ceil(np.sqrt(2 * 256 * world)))                       # >= 2 full decode batches on every rank
    ma.setup_grid(samples=warm, padding=0.1)
    ma.scan_scores(targets=targets)                                     # warm-up: cuDNN algorithm choice, allocator, NCCL
    ma.setup_grid(samples=G, padding=0.1)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rows = ma.scan_scores(targets=targets)                              # ends with the gather + D2H of [G^2, 20]
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    checksum = rows_checksum(rows)
    rc_rel, rc_n = recheck(ma, rows, 256, targets=targets)
    fp32_decode = None
    if world == 1:
        # distance of the TF32-decoded surfaces (torch's default convolution math, the headline) from an fp32 decode, on
        # 1024 grid points spread over the grid (four decode batches; cuDNN's fp32 path is ~50x slower than TF32 here)
        idx = torch.linspace(0, G * G - 1, 1024).long()
        old_tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            ma.scan_scores_at(idx[:256], targets=targets)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            rows32 = ma.scan_scores_at(idx, targets=targets)
            f1.record()
            torch.cuda.synchronize()
        finally:
            torch.backends.cudnn.allow_tf32 = old_tf32
        d = ((rows[idx].double() - rows32.double()).abs() / rows32.double().abs().clamp_min(1e-30))
        fp32_decode = {"points": int(len(idx)), "value": len(idx) / (f0.elapsed_time(f1) * 1e-3), "unit": "points/s",
                       "tf32_vs_fp32_rel_median": {"energy": float(d[:, :4].median()), "rmsd": float(d[:, 4:].median())},
                       "tf32_vs_fp32_rel_max": {"energy": float(d[:, :4].max()), "rmsd": float(d[:, 4:].max())},
                       "note": "the headline scan decodes with torch's default TF32 convolution math (the reference's decode does "
                               "too); an untrained decoder emits collapsed structures whose energies move by per cent under TF32-sized "
                               "coordinate noise, the RMSD columns by 1e-6; the kernels' parity is checked on the decoded structures "
                               "themselves (tests/test_scan_gpu.py, TF32 on and off, all 36 points)"}
    plain = None
    if world == 1:
        # the same scan decoding through network.decode as the reference writes it (Conv1d -> BatchNorm1d -> ReLU, NCHW)
        ma.set_network(net, fuse_decoder=False)
        ma.setup_grid(samples=warm, padding=0.1)
        ma.scan_scores(targets=targets)
        ma.setup_grid(samples=G, padding=0.1)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        ma.scan_scores(targets=targets)
        p1.record()
        torch.cuda.synchronize()
        plain = {"value": G * G / (p0.elapsed_time(p1) * 1e-3), "unit": "points/s", "ms": p0.elapsed_time(p1),
                 "mode": "network.decode unfused (separate BatchNorm, bias and ReLU kernels, NCHW)"}
        ma.set_network(net)
    # share of the energy + RMSD kernels: the same batches without the decoder
    tpe = ma.physics()
    xb = torch.randn(256, frames.shape[1], 3, device=dev) * 5 + torch.from_numpy(frames[0]).to(dev)
    tg = targets.to(dev)
    from molearn_b200.analysis import rmsd_to_targets
    n_batches = -(-ma.scan_stats["points_this_rank"] // 256)
    for _ in range(2):
        tpe.energy_per_conformation(xb.permute(0, 2, 1))
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(n_batches):
        tpe.energy_per_conformation(xb.permute(0, 2, 1))
        rmsd_to_targets(xb, tg)
    s1.record()
    torch.cuda.synchronize()
    score_ms = s0.elapsed_time(s1)
    cpu = None
    if rank == 0 and world == 1 and not getattr(run_scan, "skip_cpu", False):
        # the reference's way, on the host cores, for a bounded sample of 16 grid points: decode on the CPU (same
        # network), PyTorch CPU port of the energy, NumPy RMSD - scaled to points/s
        from oracle.torch_port import TorchCPUPort
        torch.set_num_threads(os.cpu_count() or 1)
        net_cpu = AutoEncoder(out_points=frames.shape[1])
        net_cpu.load_state_dict(net.state_dict())
        net_cpu.eval()
        port = TorchCPUPort(ma.physics().topology.as_dict())
        zs = ma._encoded["grid"].reshape(-1, 2)[:: (G * G) // 16][:16]
        tgt = torch.from_numpy(frames)
        c0 = time.perf_counter()
        with torch.no_grad():
            dec = net_cpu.decode(zs)[:, :frames.shape[1], :] * data.std + data.mean
            xr = dec.permute(0, 2, 1).contiguous()
            e_cpu = torch.stack(list(port.bonded(xr)) + [port.nonbonded(xr)])
            r_cpu = ((dec.unsqueeze(1) - tgt.unsqueeze(0)) ** 2).sum(-1).mean(-1).sqrt()
        c1 = time.perf_counter() - c0
        cpu = {"value": 16 / c1, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"16 of the {G * G} grid points: CPU decode + PyTorch CPU port energy (forward) + NumPy-style RMSD, {c1:.2f} s"}
    return {"metric": "latent-grid points/sec (MurD, decode + 4 energy terms + RMSD to 16 frames)", "grid": G, "cpu_baseline": cpu,
            "value": G * G / (ms * 1e-3), "unit": "points/s", "ms": ms, "scaling": "strong", "n_gpus": world,
            "checksum": checksum, "recheck_max_rel": rc_rel, "recheck_points": rc_n, "fp32_decode": fp32_decode,
            "columns": int(rows.shape[1]), "gather_bytes": ma.scan_stats["gather_bytes"],
            "points_this_rank": ma.scan_stats["points_this_rank"],
            "energy_rmsd_kernels_ms_this_rank": score_ms, "energy_rmsd_share": score_ms / ms,
            "decoder": "FoldingNet AutoEncoder(out_points=2145), random init seed 0, evaluated by models.FusedDecoder (eval-mode BatchNorm "
                       "folded, cuDNN conv+bias+relu, channels-last; torch default TF32 conv math)",
            "plain_decode": plain,
            "wall_s": time.perf_counter() - t0}


def run_config5(dev, rank, world, info, frames, barrier, G, cpu_baseline=False):
    """BASELINE config 5: 24 MurD tiles on a 3x3x3 lattice (N = 51 432), FoldingNet decoder of that size,
    setup_grid(G, bounds=(-1,1,-1,1)), 4 energy terms per point, sharded + gathered.  G = 128 on 8 GPUs (the
    configuration BASELINE.json names); fewer GPUs run a smaller grid of the same system so that the default bench
    stays within minutes."""
    import torch.distributed as dist
    from molearn_b200.analysis import MolearnAnalysis
    from molearn_b200.data import CoordinateSet, tile_system
    from molearn_b200.models import AutoEncoder
    tinfo, x = tile_system(info, frames, 24, batch=1)
    data = CoordinateSet(tinfo, x).prepare_dataset()
    torch.manual_seed(0)
    net = AutoEncoder(out_points=x.shape[1]).to(dev)
    ma = MolearnAnalysis(batch_size=8)
    ma.set_network(net)
    ma.set_dataset("complex", data)
    ma.setup_grid(samples=int(np.ceil(np.sqrt(2 * 8 * world))), bounds=(-1, 1, -1, 1), padding=0.1)
    ma.scan_scores()                                                    # warm-up: >= 2 full decode batches per rank
    ma.setup_grid(samples=G, bounds=(-1, 1, -1, 1), padding=0.1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rows = ma.scan_scores()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    rc_rel, rc_n = recheck(ma, rows, 16)
    dv = ma.physics()._dev()
    ws1 = int(_lib_ws(dv, 8)) // 8                                    # at the scan's decode batch of 8
    cpu = None
    if cpu_baseline and rank == 0 and world == 1:
        # The reference cannot even be constructed at this size on the host (five dense [N,N] fp32 matrices = 53 GB,
        # BASELINE.md row 5).  Scaled figure: CPU decode of ONE grid point with the full-size network (measured) + the
        # PyTorch CPU port's energy (forward) at 2 tiles (N = 4286), scaled by (N/4286)^2 - stated as scaled.
        from oracle.torch_port import TorchCPUPort
        torch.set_num_threads(os.cpu_count() or 1)
        net_cpu = AutoEncoder(out_points=x.shape[1])
        net_cpu.load_state_dict(net.state_dict())
        net_cpu.eval()
        z1 = ma._encoded["grid"].reshape(-1, 2)[:1]
        c0 = time.perf_counter()
        with torch.no_grad():
            net_cpu.decode(z1)
        t_dec = time.perf_counter() - c0
        sinfo, sx = tile_system(info, frames, 2, batch=1)
        from molearn_b200 import Topology
        port = TorchCPUPort(Topology(sinfo).as_dict())
        xs = torch.from_numpy(sx).permute(0, 2, 1).contiguous()
        with torch.no_grad():
            port.bonded(xs); port.nonbonded(xs)
            c0 = time.perf_counter()
            port.bonded(xs); port.nonbonded(xs)
            t_e = time.perf_counter() - c0
        scale = (x.shape[1] / sx.shape[1]) ** 2
        cpu = {"value": 1.0 / (t_dec + t_e * scale), "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port", "scaled": True,
               "sample": f"1 grid point: CPU decode at N = {x.shape[1]} ({t_dec:.2f} s, measured) + CPU port energy forward at N = {sx.shape[1]} "
                         f"({t_e:.3f} s) scaled x{scale:.0f} (N^2) - the reference cannot be built at this N (53 GB of dense tables)"}
    return {"n_atoms": int(x.shape[1]), "grid": G, "value": G * G / (ms * 1e-3), "unit": "points/s", "ms": ms,
            "n_gpus": world, "columns": int(rows.shape[1]), "gather_bytes": ma.scan_stats["gather_bytes"],
            "checksum": rows_checksum(rows), "nonfinite_entries": rows_nonfinite(rows), "recheck_max_rel": rc_rel, "recheck_points": rc_n,
            "finite_rows": int(torch.isfinite(rows).all(dim=1).sum()), "workspace_bytes_per_conformation": ws1,
            "workspace_bytes_single_conformation_call": int(_lib_ws(dv, 1)),
            "cpu_baseline": cpu, "info": dv.info()}


def _lib_ws(dv, B):
    from molearn_b200 import _lib
    return _lib.lib().mlp_energy_workspace_bytes(dv._h, B, _lib.TERM_ALL, 1)


def run_config2(dev, info, frames):
    """BASELINE config 2: one Torch_Physics_Trainer training step, MurD, batch 64 (FoldingNet AE forward/backward,
    physics loss = bonded terms of 32 interpolated decodes, AdamW update).  Reports the step time and the share of
    the physics kernels (forward + backward correction pass), timed by CUDA events inside the C ABI."""
    from molearn_b200.data import CoordinateSet
    from molearn_b200.models import AutoEncoder
    from molearn_b200.trainers import Torch_Physics_Trainer
    data = CoordinateSet(info, frames).prepare_dataset()
    torch.manual_seed(0)
    tr = Torch_Physics_Trainer(device=dev)
    tr.set_data(data)
    tr.set_autoencoder(AutoEncoder, out_points=frames.shape[1])
    tr.prepare_optimiser()
    tr.prepare_physics(physics_scaling_factor=0.1)
    batch = ((make_batch(frames, 0) - data.mean) / data.std).to(dev)
    for _ in range(3):
        tr.training_iteration(batch)
    torch.cuda.synchronize()
    dv = tr.physics_loss._dev()
    dv.profile(True)
    K = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        res = tr.training_iteration(batch)
    e1.record()
    torch.cuda.synchronize()
    pr = dv.profile_read()
    dv.profile(False)
    ms = e0.elapsed_time(e1) / K
    phys_ms = pr["bonded"][0] / K
    # beside it: the physics loss of the same 32 decoded conformations in the reference's formulation (roll groups,
    # bonded terms only, fwd + bwd: what torch_physics_trainer.py:24-45 calls) on this GPU and on the host cores
    base = {}
    if not getattr(run_config2, "skip_baseline", False):
        gen = (tr._internal["generated"].detach() * tr.std + tr.mean).float()          # [32,N,3] Angstrom
        topo = tr.physics_loss.topology.as_dict()
        base["torch_cuda"] = torch_cuda_rate(topo, gen.cpu(), dev, reps=5, nonbonded=False)
        from oracle.torch_port import TorchCPUPort
        torch.set_num_threads(os.cpu_count() or 1)
        port = TorchCPUPort(topo)
        xc = gen.cpu().permute(0, 2, 1).contiguous()
        port.energy_and_grad(xc, nonbonded=False)
        c0 = time.perf_counter()
        port.energy_and_grad(xc, nonbonded=False)
        c1 = time.perf_counter() - c0
        base["cpu"] = {"ms_per_step": 1e3 * c1, "cores": torch.get_num_threads(), "kind": "port",
                       "sample": "bonded terms of the 32 generated conformations, fwd+bwd, PyTorch CPU port (roll formulation)"}
    return {"ms_per_step": ms, "batch": B, "physics_conformations": B // 2, "physics_kernels_ms_per_step": phys_ms,
            "physics_loss_baselines": base,
            "physics_share": phys_ms / ms, "bonded_launches_per_step": pr["bonded"][1] / K,
            "loss": float(res["loss"]), "physics_loss": float(res["physics_loss"]),
            "note": "step time is dominated by the FoldingNet autoencoder (cuDNN); the physics loss is 2 launches"}


def run_config4(dev, info, frames, pk):
    """BASELINE config 4: synthetic 10 715-atom system (5 MurD tiles), all four terms + gradient, B = 128."""
    from molearn_b200 import TorchProteinEnergy, _lib
    from molearn_b200.data import tile_system
    BB = 128
    tinfo, x = tile_system(info, frames, 5, batch=BB, noise=0.05, seed=1234)
    tpe = TorchProteinEnergy(None, tinfo, device=dev)
    dv = tpe._dev()
    xd = torch.from_numpy(x).to(dev)
    g = torch.empty_like(xd)
    e = torch.empty((BB, 4), device=dev)
    for _ in range(2):
        dv.eval(xd.permute(0, 2, 1), e, g.permute(0, 2, 1), None, _lib.TERM_ALL)
    torch.cuda.synchronize()
    dv.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        dv.eval(xd.permute(0, 2, 1), e, g.permute(0, 2, 1), None, _lib.TERM_ALL)
    e1.record()
    torch.cuda.synchronize()
    pr = dv.profile_read()
    dv.profile(False)
    ms = e0.elapsed_time(e1) / K
    nb_ms = pr["nb"][0] / pr["nb"][1]
    flops = FLOPS_PER_PAIR * tpe.topology.n_nb_pairs * BB
    peak = 148 * 128 * 2 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    cpu = None
    try:   # the reference cannot hold [128,N,N]: one conformation at a time (three dense [N,N] fp32 tables + autograd, ~7 GB)
        from oracle.torch_port import TorchCPUPort
        torch.set_num_threads(os.cpu_count() or 1)
        port = TorchCPUPort(tpe.topology.as_dict())
        x1 = torch.from_numpy(x[:1]).permute(0, 2, 1).contiguous()
        port.energy_and_grad(x1)
        c0 = time.perf_counter()
        port.energy_and_grad(x1)
        c1 = time.perf_counter() - c0
        cpu = {"value": 1.0 / c1, "unit": "conformations/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"1 of the {BB} conformations, fwd+bwd of all four terms, {c1:.2f} s"}
        del port
    except Exception as exc:                      # e.g. not enough host memory
        cpu = {"unavailable": str(exc)[:200]}
    tc = None
    if cpu is not None and "unavailable" not in cpu:
        del g, e
        torch.cuda.empty_cache()
        try:   # the reference's dense [B,N,N] chain on this GPU: 0.46 GB per conformation and intermediate -> a batch of 8
            tc = torch_cuda_rate(tpe.topology.as_dict(), torch.from_numpy(x[:8]), dev, reps=2)
        except Exception as exc:
            tc = {"unavailable": str(exc)[:200]}
    return {"n_atoms": tpe.n_atoms, "batch": BB, "conformations_per_s": BB / (ms * 1e-3), "ms_per_step": ms, "cpu_baseline": cpu,
            "torch_cuda_baseline": tc, "workspace_bytes": int(_lib_ws(dv, BB)),
            "pairs_per_s": tpe.topology.n_nb_pairs * BB / (nb_ms * 1e-3), "nb_kernel_ms": nb_ms,
            "nb_fp32_frac": flops / (nb_ms * 1e-3) / 1e12 / peak, "bonded_kernel_ms": pr["bonded"][0] / pr["bonded"][1],
            "nb_reduce_ms": pr["nb_reduce"][0] / pr["nb_reduce"][1], "info": dv.info(),
            "note": "input (16.5 MB) < L2; the NB kernel is FMA-bound, its memory traffic is the partial-force scratch (workspace_bytes)"}


def run_heavy(dev, pk):
    """All-heavy-atom MurD (N = 3286; the reference's worst case: 155 roll groups), batch 64, four terms +
    gradient: whole-call rate and per-kernel times."""
    from molearn_b200 import TorchProteinEnergy, _lib
    d = np.load(os.path.join(ROOT, "tests", "golden", "murd_input.npz"))
    info = np.empty((len(d["names"]), 3), dtype=object)
    info[:, 0], info[:, 1], info[:, 2] = [str(a) for a in d["names"]], [str(a) for a in d["resnames"]], [int(r) for r in d["resids"]]
    tpe = TorchProteinEnergy(None, info, device=dev)
    dv = tpe._dev()
    x = make_batch(np.ascontiguousarray(d["coords"]), 3).to(dev)
    g = torch.empty_like(x)
    e = torch.empty((B, 4), device=dev)
    for _ in range(3):
        dv.eval(x.permute(0, 2, 1), e, g.permute(0, 2, 1), None, _lib.TERM_ALL)
    torch.cuda.synchronize()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        dv.eval(x.permute(0, 2, 1), e, g.permute(0, 2, 1), None, _lib.TERM_ALL)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    dv.profile(True)
    for _ in range(K):
        dv.eval(x.permute(0, 2, 1), e, g.permute(0, 2, 1), None, _lib.TERM_ALL)
    pr = dv.profile_read()
    dv.profile(False)
    nb_ms = pr["nb"][0] / pr["nb"][1]
    peak = 148 * 128 * 2 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    top = tpe.topology
    return {"n_atoms": tpe.n_atoms, "batch": B, "conformations_per_s": B / (ms * 1e-3), "ms_per_step": ms,
            "terms": {"bonds": len(top.bond_idxs), "angles": len(top.angle_idxs), "torsions": len(top.torsion_idxs)},
            "nb_kernel_ms": nb_ms, "nb_fp32_frac": FLOPS_PER_PAIR * top.n_nb_pairs * B / (nb_ms * 1e-3) / 1e12 / peak,
            "bonded_kernel_ms": pr["bonded"][0] / pr["bonded"][1], "nb_reduce_ms": pr["nb_reduce"][0] / pr["nb_reduce"][1],
            "launches_per_step": 3, "info": dv.info()}


def fp32_probe(dev):
    """Live FP32 FMA-pipe peak of this GPU: the library's probe kernel (mlp_probe_fp32), FFMA and FFMA2, CUDA events."""
    from molearn_b200 import _lib
    L = _lib.lib()
    n = L.mlp_probe_fp32(None, 0, 0, 1, None)
    buf = torch.empty(n, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    out = {}
    for packed, name in ((0, "ffma"), (1, "ffma2")):
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            flops = L.mlp_probe_fp32(buf.data_ptr(), n, packed, 20000, st)
            e1.record()
            torch.cuda.synchronize()
            if flops > 0:
                best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out[name + "_tflops"] = best
    return out


def ncu_summary():
    """Counters that only a profiler sees (DRAM bytes, pipe utilisation) come from the committed summary of the last
    capture, with the commit it was taken at - never from constants in this file."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_summary.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline_bonded / scan / large-N extras")
    ap.add_argument("--scan-grid", type=int, default=256, help="G of the GxG latent grid of the scan section")
    ap.add_argument("--bonded-only", action="store_true", help="tuning aid: only the headline + roofline_bonded sections")
    ap.add_argument("--config5", action="store_true", help="run BASELINE config 5 (51 432-atom scan) even with --no-extras")
    ap.add_argument("--config5-grid", type=int, default=0, help="G of the config-5 grid (default: 128 on 8 GPUs, 16*sqrt(N) below)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    from molearn_b200 import TorchProteinEnergy, _lib
    info, frames = murd()
    tpe = TorchProteinEnergy(None, info, device=dev)
    top = tpe.topology
    N = top.n_atoms
    dv = tpe._dev()
    pk, pk_kind = peaks()

    # pool of distinct batches > L2 (126 MB): P * 64 * N * 12 B
    per = B * N * 12
    P = max(8, int(np.ceil(160e6 / per)))
    xs = torch.stack([make_batch(frames, 1000 * rank + i) for i in range(P)]).to(dev)     # [P,B,N,3]
    gs = torch.empty_like(xs)
    en = torch.empty((P, B, 4), device=dev)

    streams = []        # filled for the "two streams" leg only; the headline uses torch's current stream

    def step(i):
        k = i % P
        if streams:
            with torch.cuda.stream(streams[i % len(streams)]):
                dv.eval(xs[k].permute(0, 2, 1), en[k], gs[k].permute(0, 2, 1), None, _lib.TERM_ALL)
        else:
            dv.eval(xs[k].permute(0, 2, 1), en[k], gs[k].permute(0, 2, 1), None, _lib.TERM_ALL)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for i in range(warm):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        for i in range(steps):
            fn(warm + i)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    l0 = dv.launches
    step(0)
    launches = (dv.launches - l0) * args.steps          # kernels of ours launched inside the timed region
    clocks = ClockSampler(local)
    clocks.start()
    ms = timed(step, args.steps, args.warmup)
    clk = clocks.stop()
    value = world * B * args.steps / (ms * 1e-3)
    # same K steps, batches alternating over two CUDA streams (independent batches, e.g. a scoring loop): the pair
    # kernel of batch i+1 fills the last wave of batch i and overlaps its force reduction.  Reported beside the
    # headline, which stays the one-batch-at-a-time number of a training loop.
    streams.extend(torch.cuda.Stream(dev) for _ in range(2))
    ms2 = timed(step, args.steps, args.warmup)
    streams.clear()

    # ---- e2e: public API, pinned HOST buffers; H2D of every batch and D2H of its energies AND gradient are
    # inside the timed region.  Two legs: "serial" (copy in -> compute -> copy out -> wait, one batch at a
    # time) and the headline "pipelined" one (the loader pattern a training / scoring loop uses: the next
    # batch's H2D and the previous batch's D2H run on side streams while the current batch computes).
    hx = [make_batch(frames, 5000 + 1000 * rank + i).pin_memory() for i in range(int(os.environ.get("MLP_E2E_SLOTS", 12)))]
    hg = [torch.empty((B, N, 3)).pin_memory() for _ in range(2)]
    he = [torch.empty((B, 4)).pin_memory() for _ in range(2)]
    dxs = [torch.empty((B, N, 3), device=dev) for _ in range(2)]

    def e2e_serial(i):
        dx = dxs[0]
        dx.copy_(hx[i % len(hx)], non_blocking=True)
        x = dx.permute(0, 2, 1).detach().requires_grad_(True)
        e = tpe.energy_per_conformation(x)
        e.sum().backward()
        he[0].copy_(e.detach(), non_blocking=True)
        hg[0].copy_(x.grad.permute(0, 2, 1), non_blocking=True)
        torch.cuda.current_stream().synchronize()                       # the caller reads the result

    e2e_steps = max(10, args.steps // 4)
    ms_ser = timed(e2e_serial, e2e_steps, 3)

    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main_s = torch.cuda.current_stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    keep = {}

    def stage_in(i):
        k = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[k])                                 # compute on this buffer has finished
            dxs[k].copy_(hx[i % len(hx)], non_blocking=True)
            ev_in[k].record(s_in)

    direct = {"on": False}

    def e2e_pipelined(i):
        k = i % 2
        if i == 0 or keep.get("primed") != i:
            stage_in(i)                                                 # first step of a timed run: nothing was prefetched
        main_s.wait_event(ev_in[k])
        if direct["on"]:                                                # explicit forces, no autograd graph
            e, g = tpe.energy_and_gradient(dxs[k].permute(0, 2, 1))
        else:                                                           # the reference-style call: energies, then backward
            x = dxs[k].permute(0, 2, 1).detach().requires_grad_(True)
            e = tpe.energy_per_conformation(x)
            e.sum().backward()
            g = x.grad
        ev_free[k].record(main_s)
        stage_in(i + 1)                                                 # H2D of the next batch overlaps what follows
        keep["primed"] = i + 1
        ev_out[k].synchronize()                                         # host buffers of step i-2 have been consumed
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_free[k])
            he[k].copy_(e.detach(), non_blocking=True)
            hg[k].copy_(g.permute(0, 2, 1), non_blocking=True)
            ev_out[k].record(s_out)
        keep[k] = (e, g)                                                # keep device tensors alive until copied out

    for k in range(2):
        ev_free[k].record(main_s)
        ev_out[k].record(s_out)

    def timed_pipelined(steps, warm):
        for i in range(warm):
            e2e_pipelined(i)
        s_out.synchronize(); s_in.synchronize()
        barrier()
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        for i in range(steps):
            e2e_pipelined(warm + i)
        s_out.synchronize(); s_in.synchronize(); main_s.synchronize()   # every result is in host memory
        e1.record(main_s)
        barrier()
        ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0) * 0.0)
        wall = 1e3 * (time.perf_counter() - t0)
        ms = max(ms, 0.0)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, wall

    pipe_steps = max(20, args.steps // 2)
    ms_pipe, wall_pipe = timed_pipelined(pipe_steps, 4)
    direct["on"] = True
    ms_dir, _ = timed_pipelined(pipe_steps, 4)

    # The same step - H2D of the batch, energy_per_conformation, backward of the energy sum, D2H of energies and
    # gradient - captured ONCE per pinned input buffer into a CUDA graph (molearn_b200.loss_functions.GraphedEnergyStep,
    # through the public autograd path) and replayed: one cudaGraphLaunch per batch instead of ~0.14 ms of Python,
    # twelve pinned buffers, each with its own stream (measured: 4 streams 379 k, 8 streams 390 k, 12 streams 422 k, 16 and 24
    # streams 400 k conformations/s - consecutive batches overlap like the two-stream device-resident mode).  The host
    # reads every batch's results (result(k) waits for batch k).
    from molearn_b200.loss_functions import GraphedEnergyStep
    prio = int(os.environ.get("MLP_GRAPH_PRIORITIES", 2))     # two stream priority levels: steady rate from the first batches on (see GraphedEnergyStep)
    gstep = GraphedEnergyStep(tpe, hx, n_streams=int(os.environ.get("MLP_GRAPH_STREAMS", 12)), priority_levels=prio)
    n_slots = len(hx)

    def timed_graph(steps, warm):
        for i in range(warm):
            gstep.submit(i % n_slots)
        gstep.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        for st in gstep.streams:
            st.wait_event(e0)
        acc = 0.0
        for i in range(steps):
            k = i % n_slots
            if i >= n_slots:
                acc += float(gstep.result(k)[0][0, 0])                  # the host consumes batch i - n_slots before its buffers are reused
            gstep.submit(k)
        for k in range(min(steps, n_slots)):
            acc += float(gstep.result(k)[0][0, 0])
        for st in gstep.streams:
            main_s.wait_stream(st)
        e1.record(main_s)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, acc

    ms_graph, _ = timed_graph(args.steps, 2 * n_slots)
    # the same loop with the gradient left on the device (a consumer on the GPU, e.g. the decoder's backward pass):
    # H2D of every batch and D2H of its energies only - half the PCIe / host-memory traffic
    gstep_full = gstep
    gstep = GraphedEnergyStep(tpe, hx, n_streams=len(gstep_full.streams), gradient_to_host=False, priority_levels=prio)
    ms_graph_dev, _ = timed_graph(args.steps, 2 * n_slots)
    gstep = gstep_full
    # parity of the replayed step with the eager call on the same host batch (bit for bit)
    ge, gg = gstep.result(0)
    xe = hx[0].to(dev).permute(0, 2, 1).requires_grad_(True)
    ee = tpe.energy_per_conformation(xe)
    ee.sum().backward()
    graph_equal = bool(torch.equal(ge, ee.detach().cpu()) and torch.equal(gg, xe.grad.permute(0, 2, 1).cpu()))
    eager = {"value": world * B * pipe_steps / (ms_pipe * 1e-3), "ms_per_step": ms_pipe / pipe_steps,
             "mode": "eager Python loop, pipelined: H2D of batch i+1 and D2H of batch i-1 (energies + gradient) overlap the kernels of "
                     "batch i; host-bound (autograd bookkeeping + launches + copies per batch on one core)"}
    e2e = {"value": world * B * args.steps / (ms_graph * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": per, "d2h_bytes_per_step": per + B * 16, "ms_per_step": ms_graph / args.steps,
           "mode": "GraphedEnergyStep: H2D + energy_per_conformation + backward + D2H of energies and gradient captured once per pinned "
                   "host buffer through the public autograd path, one CUDA-graph replay per batch, " + f"{n_slots} pinned buffers over {len(gstep.streams)} streams with {prio} priority levels (consecutive batches overlap on the GPU like the two_streams mode); the host reads every result",
           "graph_equals_eager": graph_equal,
           "gradient_on_device": {"value": world * B * args.steps / (ms_graph_dev * 1e-3), "ms_per_step": ms_graph_dev / args.steps,
                                  "h2d_bytes_per_step": per, "d2h_bytes_per_step": B * 16,
                                  "mode": "same graph replay, dE/dx stays on the device (only the energies [B,4] are read back): what a training "
                                          "step does - the gradient feeds the decoder's backward pass.  On 8 GPUs this leg scales 8.0x while the full "
                                          "read-back scales 6.4x: eight ranks reading 1.6 MB of gradient per 0.15 ms saturate the host's memory system, not the GPUs"},
           "eager_pipelined": eager,
           "pipelined_energy_and_gradient_call": {"value": world * B * pipe_steps / (ms_dir * 1e-3), "ms_per_step": ms_dir / pipe_steps,
                                                  "mode": "same eager pipeline through TorchProteinEnergy.energy_and_gradient (no autograd graph)"},
           "serial": {"value": world * B * e2e_steps / (ms_ser * 1e-3), "ms_per_step": ms_ser / e2e_steps,
                      "mode": "copy in -> compute -> copy out -> host wait, one batch at a time"}}
    del gstep

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(N), "clocks": clk, "e2e": e2e,
        "gpu_launches": int(launches),
        "two_streams": {"value": world * B * args.steps / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / args.steps,
                        "mode": "same K steps, independent batches alternating over two CUDA streams (one handle)"},
    }

    # ---- latent-grid scan (BASELINE config 3): sharded over the ranks, one all-gather --------------
    run_scan.skip_cpu = args.no_cpu_baseline
    if not args.no_extras and not args.bonded_only:
        out["scan"] = run_scan(dev, rank, world, info, frames, args.scan_grid, barrier)

    # ---- BASELINE config 5 (51 432 atoms): the named 128 x 128 grid on 8 GPUs; the same system on a smaller grid
    # with fewer GPUs, so that every default run exercises it ------------------------------------------------------
    c5_grid = args.config5_grid or (128 if world >= 8 else 16 * max(1, int(round(np.sqrt(world)))))
    if (not args.no_extras and not args.bonded_only) or args.config5:
        out["config5_50k_scan"] = run_config5(dev, rank, world, info, frames, barrier, c5_grid,
                                              cpu_baseline=not args.no_cpu_baseline)

    if rank == 0:
        # ---- per-kernel durations (CUDA events on the launching stream, inside the C ABI) ----
        dv.profile(True)
        for i in range(max(3, args.warmup)):                 # warm-up of the per-kernel timing mode (it runs the bonded kernel on the main stream)
            step(i)
        dv.profile_read()
        for i in range(args.steps):
            step(i)
        prof = dv.profile_read()
        dv.profile(False)
        nb_ms = prof["nb"][0] / max(prof["nb"][1], 1)
        bd_ms = prof["bonded"][0] / max(prof["bonded"][1], 1)
        red_ms = prof["nb_reduce"][0] / max(prof["nb_reduce"][1], 1)
        fp32_peak = 148 * 128 * 2 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12       # TFLOP/s, FMA pipe
        flops = FLOPS_PER_PAIR * top.n_nb_pairs * B
        ach = flops / (nb_ms * 1e-3) / 1e12
        ncu = ncu_summary()
        probe = fp32_probe(dev)
        out["roofline"] = {"kernel": "non-bonded pair kernel <grad> (packed FP32, FFMA2)", "bound": "fp32", "achieved": ach, "peak": fp32_peak,
                           "unit": "TFLOP/s", "frac": ach / fp32_peak,
                           "traffic": ncu.get("nb", {}).get("dram_bytes"), "ncu": ncu.get("nb"), "ncu_capture_commit": ncu.get("commit"),
                           "measured_fp32_peak": probe, "frac_of_measured_ffma2_peak": ach / probe["ffma2_tflops"] if probe.get("ffma2_tflops") else None,
                           "peak_source": "148 SM x 128 FMA lanes x 2 x sm_max_mhz (FP32 pipe; no tensor cores on this path; MEASURED_PEAKS.json "
                                          "holds no FP32 figure, so the nominal pipe rate is the denominator and the probe's live rate is beside it)",
                           "algorithmic_flops_per_launch": flops, "kernel_ms": nb_ms,
                           "share_of_step": nb_ms / (ms / args.steps),
                           "other_kernels_ms": {"bonded": bd_ms, "nb_reduce": red_ms},
                           "note": "per-kernel times are CUDA-event intervals on one stream (profiling mode serialises the bonded kernel, "
                                   "which otherwise overlaps the pair kernel on a side stream)"}
        if not args.no_extras:
            # bonded kernel against HBM at a batch that is not launch-latency bound
            BB = 4096
            xb = make_batch(frames, 77, 64).to(dev).repeat(BB // 64, 1, 1)
            xb += 0.01 * torch.randn_like(xb)
            gb = torch.empty_like(xb)
            eb = torch.empty((BB, 4), device=dev)
            dv.profile(True)
            for _ in range(13):
                dv.eval(xb.permute(0, 2, 1), eb, gb.permute(0, 2, 1), None, _lib.TERM_BONDED)
            pr = dv.profile_read()
            dv.profile(False)
            t_ms = pr["bonded"][0] / pr["bonded"][1]
            topo_bytes = len(top.bond_idxs) * 16 + len(top.angle_idxs) * 32 + len(top.torsion_idxs) * 64
            byts = 24 * N * BB + topo_bytes
            gbs = byts / (t_ms * 1e-3) / 1e9
            nb_ncu = ncu.get("bonded_b4096", {})       # the capture of this configuration (tools/ncu_round2.sh)
            out["roofline_bonded"] = {"kernel": "bonded_kernel<grad>", "bound": "hbm", "batch": BB, "achieved": gbs,
                                      "peak": pk["hbm_gbs"], "peak_kind": pk_kind, "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                                      "traffic": nb_ncu.get("dram_bytes"), "ncu": nb_ncu or None, "ncu_capture_commit": ncu.get("commit"),
                                      "issue_frac": (nb_ncu.get("issue_active_pct") or 0) / 100 or None,
                                      "real_bound": "instruction issue / latency (see issue_frac: executed warp-instructions per issue slot, "
                                                    "from the ncu capture); the HBM fraction is what the spec's accounting asks for",
                                      "algorithmic_bytes_per_launch": byts, "kernel_ms": t_ms,
                                      "conformations_per_s": BB / (t_ms * 1e-3)}
            del xb, gb, eb
        if world == 1 and not args.no_extras and not args.bonded_only:
            run_config2.skip_baseline = args.no_cpu_baseline
            out["config2_training_step"] = run_config2(dev, info, frames)
            out["config4_10k_atoms"] = run_config4(dev, info, frames, pk)
            out["murd_all_heavy_atoms"] = run_heavy(dev, pk)
        if world == 1 and not args.no_cpu_baseline:
            rate, sec, cores = cpu_port_rate(top.as_dict(), make_batch(frames, 0)[:CPU_SAMPLE], reps=5)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{CPU_SAMPLE} of the {B} conformations, fwd+bwd of all four terms, "
                                             f"median of 5 ({sec:.3f} s each), PyTorch CPU port of the reference (roll + cdist)"}
            tc = torch_cuda_rate(top.as_dict(), make_batch(frames, 0), dev, reps=5)
            tc["ours_over_torch_cuda"] = value / world / tc["value"]
            tc["our_kernel_launches_per_step"] = int(launches // max(args.steps, 1))
            out["torch_cuda_baseline"] = tc
        # ---- short flat keys LAST: a tail of the line still carries every headline number -------------------
        tailkeys = {"n_gpus": world, "value": value, "e2e_value": e2e["value"], "e2e_grad_on_device_value": e2e["gradient_on_device"]["value"], "e2e_eager_value": e2e["eager_pipelined"]["value"],
                    "roofline_frac": out["roofline"]["frac"]}
        if "scan" in out:
            sc = out["scan"]
            tailkeys.update(scan_pts_s=sc["value"], scan_ms=sc["ms"], scan_checksum=sc["checksum"], scan_recheck_max_rel=sc["recheck_max_rel"])
        if "config5_50k_scan" in out:
            c5 = out["config5_50k_scan"]
            tailkeys.update(config5_pts_s=c5["value"], config5_ms=c5["ms"], config5_grid=c5["grid"], config5_checksum=c5["checksum"],
                            config5_nonfinite=c5["nonfinite_entries"], config5_recheck_max_rel=c5["recheck_max_rel"])
        if "torch_cuda_baseline" in out:
            tailkeys["ours_over_torch_cuda"] = out["torch_cuda_baseline"]["ours_over_torch_cuda"]
        out["summary"] = tailkeys
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
