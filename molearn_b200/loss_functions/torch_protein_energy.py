"""Drop-in ``TorchProteinEnergy`` backed by the sm_100a kernels (C ABI ``mlp_energy_*``).

Same constructor, method names, argument meaning and return conventions as the reference class
(src/molearn/loss_functions/torch_protein_energy.py:18-68, 197-200, 224-239, 276-305), so
``Torch_Physics_Trainer`` (src/molearn/trainers/torch_physics_trainer.py:22,35) can use it
unchanged.  Differentiability comes from a ``torch.autograd.Function`` that computes energies AND
dE/dx in one fused forward pass and scales the saved gradient in backward - the pattern of the
reference's own native energy, ``openmm_energy_function`` (openmm_thread.py:290-321).

There is no CPU path: a CPU ``device`` raises.
"""
from __future__ import annotations

import ctypes

import torch

from .. import _lib
from ..topology import Topology

TERM_BOND, TERM_ANGLE, TERM_TORSION, TERM_NB = _lib.TERM_BOND, _lib.TERM_ANGLE, _lib.TERM_TORSION, _lib.TERM_NB
TERM_BONDED, TERM_ALL = _lib.TERM_BONDED, _lib.TERM_ALL


class _DeviceEnergy:
    """Owns one ``mlp_energy`` handle (device copy of the packed topology)."""

    def __init__(self, topology: Topology, device: torch.device, nb_mode: int):
        self.topology = topology
        self.device = device
        self._h = ctypes.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.index = idx
        _lib.check(_lib.lib().mlp_energy_create(topology.handle, idx, nb_mode, ctypes.byref(self._h)))
        self.launches = 0  # kernels launched through this handle (bench.py reports it)
        self._ws_bytes = {}
        self._ws = None

    # scratch memory per call is capped: larger batches are evaluated in chunks of conformations
    MAX_BATCH_PER_CALL = 32768      # the library takes at most 65535 conformations per call; chunks stay well below
    MAX_WORKSPACE_BYTES = int(float(__import__("os").environ.get("MLP_MAX_WORKSPACE_GB", "8")) * 2 ** 30)

    def _workspace_bytes(self, B, term_mask, want_grad):
        key = (B, term_mask, want_grad)
        nbytes = self._ws_bytes.get(key)
        if nbytes is None:
            nbytes = self._ws_bytes[key] = _lib.lib().mlp_energy_workspace_bytes(self._h, B, term_mask, int(want_grad))
        return nbytes

    def _workspace(self, nbytes, device):
        """Scratch buffer of the call.  The last one is kept and handed out again: allocations made on one stream
        are reused by later calls on the same stream only (the caching allocator's rule, applied by hand)."""
        stream = torch.cuda.current_stream(device).cuda_stream
        if torch.cuda.is_current_stream_capturing():
            # CUDA-graph capture: the scratch must come from the graph's own memory pool and stay with the graph
            return torch.empty(nbytes, dtype=torch.uint8, device=device), stream
        ws = self._ws
        if ws is None or ws[0].numel() < nbytes or ws[1] != stream:
            ws = self._ws = (torch.empty(nbytes, dtype=torch.uint8, device=device), stream)
        return ws[0], stream

    def eval(self, x, energies, grad, weights, term_mask, accumulate=False):
        """x: [B,3,N] logical fp32 CUDA tensor (any strides). energies [B,4] or None; grad like x or None."""
        B = x.shape[0]
        want_grad = grad is not None
        chunk = min(B, self.MAX_BATCH_PER_CALL)
        # halve the chunk until its workspace fits (the library's sizes are not exactly linear in B: alignment slack)
        while chunk > 1 and self._workspace_bytes(chunk, term_mask, want_grad) > self.MAX_WORKSPACE_BYTES:
            chunk = (chunk + 1) // 2
        if chunk >= B:
            return self._eval_one(x, energies, grad, weights, term_mask, accumulate)
        for lo in range(0, B, chunk):
            sl = slice(lo, min(B, lo + chunk))
            self._eval_one(x[sl], None if energies is None else energies[sl], None if grad is None else grad[sl],
                           None if weights is None else weights[sl], term_mask, accumulate)

    def _eval_one(self, x, energies, grad, weights, term_mask, accumulate):
        L = _lib.lib()
        B = x.shape[0]
        nbytes = self._workspace_bytes(B, term_mask, grad is not None)
        ws, stream = self._workspace(nbytes, x.device)
        gs = grad.stride() if grad is not None else (0, 0, 0)
        xs = x.stride()
        rc = L.mlp_energy_eval(
            self._h, x.data_ptr(), B, xs[0], xs[1], xs[2],
            energies.data_ptr() if energies is not None else None,
            grad.data_ptr() if grad is not None else None, gs[0], gs[1], gs[2],
            weights.data_ptr() if weights is not None else None,
            term_mask, _lib.EVAL_ACCUMULATE_GRAD if accumulate else 0,
            ws.data_ptr(), nbytes, stream)
        if rc < 0:
            _lib.check(rc)
        self.launches += rc                                             # kernels queued by this call

    def profile(self, enable: bool):
        _lib.check(_lib.lib().mlp_energy_profile(self._h, int(enable)))

    def profile_read(self):
        """{'bonded'|'nb'|'nb_reduce': (total ms, launches)} since the last read (waits for the events)."""
        import numpy as np
        ms, n = np.zeros(4, np.float64), np.zeros(4, np.int64)
        _lib.check(_lib.lib().mlp_energy_profile_read(self._h, ms.ctypes.data_as(ctypes.c_void_p),
                                                      n.ctypes.data_as(ctypes.c_void_p)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("unused", "bonded", "nb", "nb_reduce")) if k != "unused"}

    def info(self):
        import numpy as np
        v = np.zeros(8, np.int64)
        _lib.check(_lib.lib().mlp_energy_info(self._h, v.ctypes.data_as(ctypes.c_void_p)))
        keys = ("bonded_tiles", "bonded_tile_atoms", "max_contributions_per_atom", "bonded_smem_bytes",
                "nb_tiles", "nb_tile_pairs", "nb_masked_tile_pairs", "n_atoms_padded")
        out = dict(zip(keys, (int(a) for a in v)))
        out["nb_table_classes"] = int(_lib.lib().mlp_energy_nb_table_classes(self._h))   # 0: w_ij computed per pair
        return out

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().mlp_energy_destroy(h)
            except Exception:
                pass


def _check_input(x, n_atoms, device=None):
    if not x.is_cuda:
        raise RuntimeError("molearn_b200 energy kernels need a CUDA tensor (there is no CPU fallback)")
    if x.dim() != 3 or x.shape[1] != 3:
        raise ValueError(f"expected coordinates of shape [B, 3, N], got {tuple(x.shape)}")
    if x.shape[2] != n_atoms:
        raise ValueError(f"expected {n_atoms} atoms, got {x.shape[2]}")
    if x.dtype != torch.float32:
        raise TypeError("coordinates must be float32")
    if device is not None and x.device.index != device:
        raise ValueError(f"coordinates live on cuda:{x.device.index}, the energy object on cuda:{device}")


class _EnergyFunction(torch.autograd.Function):
    """E[b, t] for t in (bond, angle, torsion, nb); terms outside ``term_mask`` are 0."""

    @staticmethod
    def forward(ctx, x, dev: _DeviceEnergy, term_mask: int):
        _check_input(x, dev.topology.n_atoms, dev.index)
        xd = x.detach()
        energies = torch.empty((x.shape[0], 4), dtype=torch.float32, device=x.device)
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(xd) if need_grad else None
        if torch.cuda.current_device() == dev.index:
            dev.eval(xd, energies, grad, None, term_mask)
        else:
            with torch.cuda.device(x.device):
                dev.eval(xd, energies, grad, None, term_mask)
        ctx.dev, ctx.term_mask = dev, term_mask
        ctx.save_for_backward(xd, grad)
        ctx.set_materialize_grads(False)
        return energies

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        if gout is None:
            return None, None, None
        x, saved = ctx.saved_tensors
        if saved is None:
            raise RuntimeError("TorchProteinEnergy: backward through a forward that did not record a gradient")
        if gout.stride() == (0, 0):
            # upstream is one scalar broadcast over [B,4] (e.g. ``E.sum().backward()``): every term of every
            # conformation has the same weight, so the saved unit-weight gradient only needs scaling
            return saved * gout.reshape(-1)[0].to(saved.dtype), None, None
        # General upstream gradient: dL/dx = sum_t gout[b,t] dE_t/dx is evaluated by the kernels with gout as per-term,
        # per-conformation weights.  A term whose weight is zero is skipped inside the kernels, so a non-finite term or
        # conformation that the loss masked out (torch.where / nansum, torch_physics_trainer.py:40-42) cannot reach
        # the network as NaN * 0; scaling the saved sum of all terms would let it through.
        gin = torch.empty_like(saved)
        w = gout.contiguous().float()
        with torch.cuda.device(x.device):
            ctx.dev.eval(x, None, gin, w, ctx.term_mask)
        return gin, None, None


class TorchProteinEnergy:
    """Amber ff14SB-style bond / angle / torsion energy plus a softened all-pairs non-bonded term.

    :param frame: example coordinates ``[3, N]`` (only its atom count is used: the v=5 topology of
        the reference is built from residue templates, not distances).
    :param pdb_atom_names: array ``[N, 3]`` of ``(atom name, residue name, resid)``.
    :param padded_residues: not supported (the reference's own path for it does not import,
        torch_protein_energy.py:165).
    :param method: ``'roll'`` or ``'convolutional'`` - two reference formulations of the same
        energy; both are served by the same explicit-list kernels here.
    :param device: a CUDA device.
    :param fix_h: CHARMM -> Amber hydrogen renames.
    :param alt_vdw: accepted for signature compatibility; the reference only reads it in a code
        path that is dead for v=5 (utils.py:527).
    :param NB: ``'repulsive'`` (``_cdist_nb``) or ``'full'`` (``_cdist_nb_full``).
    """

    def __init__(self, frame, pdb_atom_names, padded_residues=False, method='roll',
                 device=None, fix_h=False, alt_vdw=[], NB='repulsive', param_dir=None):
        if device is None:
            device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError("molearn_b200.TorchProteinEnergy runs on CUDA devices only (no CPU fallback); "
                               f"got device={device}")
        if padded_residues or method == 'indexed':
            raise NotImplementedError("padded_residues / method='indexed' is not available "
                                      "(broken in the reference too: torch_protein_energy.py:165)")
        if method not in ('roll', 'convolutional'):
            raise ValueError(f"unknown method {method!r}")
        if NB not in ('repulsive', 'full'):
            raise ValueError(f"unknown NB {NB!r}")
        self.device = device
        self.method = method
        self.NB = NB
        self.topology = Topology(pdb_atom_names, fix_h=fix_h, param_dir=param_dir)
        if frame is not None:
            n = frame.shape[1]
            assert n == self.topology.n_atoms, f"frame has {n} atoms, pdb_atom_names {self.topology.n_atoms}"
        self.n_atoms = self.topology.n_atoms
        self._devs = {}
        self._dev(NB)
        self.get_bonded_energy = self._bonded_roll_loss if method == 'roll' else self._bonded_convolutional_loss
        self._nb_loss = self._cdist_nb_full if NB == 'full' else self._cdist_nb

    # -- plumbing ---------------------------------------------------------------------------
    def _dev(self, NB=None) -> _DeviceEnergy:
        NB = NB or self.NB
        if NB not in self._devs:
            mode = _lib.NB_FULL if NB == 'full' else _lib.NB_REPULSIVE
            self._devs[NB] = _DeviceEnergy(self.topology, self.device, mode)
        return self._devs[NB]

    @property
    def kernel_launches(self) -> int:
        return sum(d.launches for d in self._devs.values())

    def _energies(self, x, term_mask, NB=None):
        return _EnergyFunction.apply(x, self._dev(NB), term_mask)

    # -- additive API -------------------------------------------------------------------------
    def energy_per_conformation(self, x, nonbonded=True):
        """``[B, 4]`` tensor of (bond, angle, torsion, nb) per conformation, differentiable in x.
        The non-bonded column is 0 when ``nonbonded=False``."""
        return self._energies(x, TERM_ALL if nonbonded else TERM_BONDED)

    def energy_and_gradient(self, x, nonbonded=True, energies=None, grad=None):
        """Energies ``[B, 4]`` and the gradient of their sum ``dE/dx`` (same logical shape as ``x``) in one
        fused pass, without building an autograd graph - the call for scoring / sampling loops that consume
        forces directly.  Optional preallocated outputs are written in place."""
        _check_input(x, self.n_atoms, self._dev().index)
        x = x.detach()
        if energies is None:
            energies = torch.empty((x.shape[0], 4), dtype=torch.float32, device=x.device)
        if grad is None:
            grad = torch.empty_like(x)
        dev = self._dev()
        if torch.cuda.current_device() == dev.index:
            dev.eval(x, energies, grad, None, TERM_ALL if nonbonded else TERM_BONDED)
        else:
            with torch.cuda.device(x.device):
                dev.eval(x, energies, grad, None, TERM_ALL if nonbonded else TERM_BONDED)
        return energies, grad

    # -- reference API ------------------------------------------------------------------------
    def get_energy(self, x, nonbonded=False):
        """torch_protein_energy.py:55-68: bonded terms are batch MEANS, the NB term a batch SUM."""
        if nonbonded:
            e = self._energies(x, TERM_ALL)
            bs = x.shape[0]
            return e[:, 0].sum() / bs, e[:, 1].sum() / bs, e[:, 2].sum() / bs, torch.nansum(e[:, 3])
        return self.get_bonded_energy(x)

    def _roll_bond_angle_torsion_loss(self, x):
        """torch_protein_energy.py:276-305: un-normalised batch sums (bond, angle, torsion)."""
        e = self._energies(x, TERM_BONDED)
        return e[:, 0].sum(), e[:, 1].sum(), e[:, 2].sum()

    def _bonded_roll_loss(self, x):
        bs = x.shape[0]
        b, a, t = self._roll_bond_angle_torsion_loss(x)
        return b / bs, a / bs, t / bs

    _bonded_convolutional_loss = _bonded_roll_loss

    def _conv_bond_loss(self, x):
        return self._energies(x, TERM_BOND)[:, 0].sum()

    def _conv_angle_loss(self, x):
        return self._energies(x, TERM_ANGLE)[:, 1].sum()

    def _conv_torsion_loss(self, x):
        return self._energies(x, TERM_TORSION)[:, 2].sum()

    def _cdist_nb(self, x, cutoff=9.0, mask=False):
        """torch_protein_energy.py:232-236 (``cutoff`` and ``mask`` are unused there too)."""
        return torch.nansum(self._energies(x, TERM_NB, 'repulsive')[:, 3])

    def _cdist_nb_full(self, x, cutoff=9.0, mask=False):
        """torch_protein_energy.py:224-230."""
        return torch.nansum(self._energies(x, TERM_NB, 'full')[:, 3])
