"""Device-side scan reducers (C ABI ``mlp_rmsd_eval``).

Replace the host NumPy / biobox loops of ``MolearnAnalysis.scan_error`` and
``scan_error_from_target`` (src/molearn/analysis/analyser.py:626-737): only ``[B, K]`` scalars
leave the GPU instead of the decoded ``[B, N, 3]`` structures.
"""
from __future__ import annotations

import torch

from .. import _lib


def rmsd_to_targets(x, targets, scale: float = 1.0, shift: float = 0.0, align: bool = False):
    """RMSD ``sqrt(mean_atoms sum_xyz (x*scale+shift - t)^2)`` of every structure to every target
    (the intended per-point formula, SURVEY S-3/S-4).  ``align=True``: after optimal rigid
    superposition (Kabsch), the quantity biobox ``Molecule.rmsd`` returns in the reference.

    :param x: ``[B, N, 3]`` float32 CUDA tensor (any strides - e.g. a slice of a decoder output).
    :param targets: ``[K, N, 3]`` float32 CUDA tensor, in the units of ``x*scale+shift``.
    :returns: ``[B, K]`` float32.
    """
    if not x.is_cuda:
        raise RuntimeError("rmsd_to_targets needs CUDA tensors (there is no CPU fallback)")
    if targets.dim() == 2:
        targets = targets.unsqueeze(0)
    B, N, _ = x.shape
    assert targets.shape[1:] == (N, 3), f"targets {tuple(targets.shape)} do not match x {tuple(x.shape)}"
    targets = targets.contiguous().float()
    K = targets.shape[0]
    out = torch.empty((B, K), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mlp_rmsd_eval(x.data_ptr(), B, N, x.stride(0), x.stride(2), x.stride(1),
                                            float(scale), float(shift), targets.data_ptr(), K, out.data_ptr(),
                                            _lib.RMSD_ALIGN if align else 0,
                                            torch.cuda.current_stream(x.device).cuda_stream))
    return out
