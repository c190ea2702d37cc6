"""``GraphedEnergyStep`` - host batch in, energies and forces out, one CUDA-graph replay per batch.

A scoring / sampling loop that feeds fixed-shape batches from pinned host memory spends more time in Python than on
the GPU once the energy of a MurD-sized batch of 64 takes 0.16 ms: the eager call costs ~0.14 ms of host time per
batch (autograd bookkeeping, three launches, copies, events).  This class captures the WHOLE step

    H2D copy of the batch  ->  TorchProteinEnergy.energy_per_conformation(x)  ->  backward of the energy sum
    ->  D2H copy of the energies [B,4] and of dE/dx [B,N,3]

once per pinned input buffer - through the public autograd path, the same ``torch.autograd.Function`` the trainer uses -
and replays it with a single ``cudaGraphLaunch`` per batch.  Buffers alternate over two CUDA streams so that the copies
of one batch overlap the kernels of the other.

    step = GraphedEnergyStep(tpe, host_batches)       # list of pinned [B, N, 3] float32 tensors (the loader's buffers)
    step.submit(k)                                    # after the loader has written batch k's buffer
    e, g = step.result(k)                             # pinned [B,4] and [B,N,3]; waits for batch k only

The reference has no counterpart (its energy is eager ATen ops); the values are those of the eager call, bit for bit.
"""
from __future__ import annotations

import torch

from .torch_protein_energy import TorchProteinEnergy


class GraphedEnergyStep:
    def __init__(self, energy: TorchProteinEnergy, host_batches, nonbonded: bool = True, n_streams: int = 2,
                 gradient_to_host: bool = True, priority_levels: int = 0):
        """``gradient_to_host=False`` keeps dE/dx on the device (``result`` then returns the device tensor): the case of
        a consumer that lives on the GPU, e.g. a decoder's backward pass - only the energies [B,4] cross PCIe."""
        self.gradient_to_host = gradient_to_host
        if not host_batches:
            raise ValueError("GraphedEnergyStep needs at least one pinned host buffer")
        self.energy = energy
        dev = energy.device
        # priority_levels = L > 1: the streams get L descending priority levels in turn.  Batches submitted together are
        # otherwise served interleaved and finish all at once (their read-backs then arrive in a burst and the host
        # re-submits in a burst); with two levels half of them are served first - a short run reaches its steady rate
        # at once (390 k instead of 358 k conformations/s over 20 batches of MurD x 64; the long-run rate of the loop with
        # the gradient read-back drops from 427 k to 396 k, the one with the gradient on the device rises to 438 k).
        lev = int(priority_levels)
        self.streams = [torch.cuda.Stream(dev, priority=(-(lev - 1) + (j % lev)) if lev > 1 else 0) for j in range(max(1, n_streams))]
        self.slots = []
        for k, hx in enumerate(host_batches):
            if hx.is_cuda or hx.dtype != torch.float32 or hx.dim() != 3 or hx.shape[1:] != (energy.n_atoms, 3):
                raise ValueError(f"host buffer {k}: expected a CPU float32 [B, {energy.n_atoms}, 3] tensor, got {tuple(hx.shape)} {hx.dtype}")
            if not hx.is_pinned():
                raise ValueError(f"host buffer {k} is not pinned (tensor.pin_memory()): the H2D copy could not be asynchronous")
            B = hx.shape[0]
            slot = {"hx": hx, "dx": torch.empty((B, energy.n_atoms, 3), device=dev),
                    "he": torch.empty((B, 4)).pin_memory(),
                    "hg": (torch.empty((B, energy.n_atoms, 3)).pin_memory() if gradient_to_host
                           else torch.empty((B, energy.n_atoms, 3), device=dev)),
                    "stream": self.streams[k % len(self.streams)], "done": torch.cuda.Event(), "graph": None}
            self.slots.append(slot)
        torch.cuda.synchronize(dev)
        for slot in self.slots:
            s = slot["stream"]
            with torch.cuda.stream(s):
                for _ in range(2):                         # warm-up outside capture (allocator, lazy initialisation)
                    self._body(slot, nonbonded)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self._body(slot, nonbonded)
            slot["graph"] = g
        torch.cuda.synchronize(dev)

    def _body(self, slot, nonbonded):
        slot["dx"].copy_(slot["hx"], non_blocking=True)
        x = slot["dx"].permute(0, 2, 1).detach().requires_grad_(True)
        e = self.energy.energy_per_conformation(x, nonbonded=nonbonded)
        (g,) = torch.autograd.grad(e.sum(), x)
        slot["he"].copy_(e.detach(), non_blocking=True)
        slot["hg"].copy_(g.permute(0, 2, 1), non_blocking=True)        # D2H, or a device copy into the slot's own buffer

    def submit(self, k: int):
        """Queue batch k (its pinned input buffer must be filled): one graph launch on the slot's stream."""
        slot = self.slots[k]
        with torch.cuda.stream(slot["stream"]):
            slot["graph"].replay()
            slot["done"].record()

    def result(self, k: int):
        """(energies [B,4], dE/dx [B,N,3]) of the last ``submit(k)``; waits for that batch.  Pinned host memory (the
        gradient is a device tensor with ``gradient_to_host=False``)."""
        slot = self.slots[k]
        slot["done"].synchronize()
        return slot["he"], slot["hg"]

    def synchronize(self):
        for s in self.streams:
            s.synchronize()

    @property
    def kernels_per_step(self) -> int:
        """Kernels of this library inside one replay (forward: bonded + pair + reduce)."""
        return 3
