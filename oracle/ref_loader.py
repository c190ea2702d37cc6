"""TEST INFRASTRUCTURE ONLY - loads the *unmodified* molearn reference as the parity oracle.

Nothing here is on the product path.  Only ``tests/``, ``oracle/make_golden.py`` and the
``cpu_baseline`` leg of ``bench.py`` may import this module, and it only works where the
reference tree exists (this container: ``/root/reference``; override with ``$MOLEARN_REF``).
The GPU box has no reference tree; tests there use the committed fixtures in
``tests/golden`` that ``oracle/make_golden.py`` produced with this loader.

The reference cannot be imported as a package here (``import molearn`` needs OpenMM,
MDAnalysis and biobox: src/molearn/__init__.py:19), so the two hot-path files are executed
under their dotted names with the in-memory shims of SURVEY.md section 8c.  No reference
source is copied into the repository:

1. ``np.bool = bool``  (removed NumPy alias used at torch_protein_energy_utils.py:749,771,793)
2. ``pkg_resources.resource_filename`` -> the directory that really holds the Amber files
   (utils.py:21,119,148 ask for ``molearn/parameters``; the files live in
   ``molearn/loss_functions/parameters``)
3. empty ``molearn`` / ``molearn.loss_functions`` packages registered in ``sys.modules``
4. NB only (``pair_exclusions=True``): the nine exclusion sites ``M[idx.T] = 0.0``
   (utils.py:818-823, 828-830) are rewritten in memory to ``M[tuple(idx.T)] = 0.0``, the
   pair indexing torch <= 1.12 performed.  Without it every row touched by a bond is zeroed
   on modern torch and E_nb is identically 0.
6. the loss accumulators ``torch.tensor(0.0).float()`` (torch_protein_energy.py:243,251,263,
   278-280) are rewritten in memory to ``torch.zeros((), dtype=x.dtype)``: identical in fp32,
   and lets the fp64 run of the oracle accumulate in fp64 (an in-place ``+=`` would otherwise
   round every group sum to fp32).
5. the local index/parameter arrays of ``get_convolutions`` are captured through
   ``sys.settrace`` (they are not returned by the reference).
"""
from __future__ import annotations

import os
import sys
import types
import contextlib
import io

import numpy as np

REF = os.environ.get("MOLEARN_REF", "/root/reference")
_LOSS_DIR = os.path.join(REF, "src", "molearn", "loss_functions")
PARAM_DIR = os.path.join(_LOSS_DIR, "parameters")

_CAPTURE_NAMES = (
    "bond_idxs", "angle_idxs", "torsion_idxs", "bond_14_idxs",
    "bond_para", "angle_para", "torsion_para",
    "atom_names", "atom_charges", "atom_R", "atom_e",
    "bond_types", "angle_types", "torsion_types",
    "vdw_14R", "vdw_14e", "q1q2_14",
)


def available() -> bool:
    return os.path.isfile(os.path.join(_LOSS_DIR, "torch_protein_energy.py"))


def _exec_module(dotted: str, path: str, patch=None):
    with open(path) as fh:
        src = fh.read()
    if patch is not None:
        src = patch(src)
    mod = types.ModuleType(dotted)
    mod.__file__ = path
    sys.modules[dotted] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def _pair_index_patch(src: str) -> str:
    n = 0
    for idx in ("bond_idxs.T", "angle_idxs[:, (0, 2)].T", "angle_idxs[:,(0, 2)].T",
                "torsion_idxs[:, (0, 3)].T", "torsion_idxs[:,(0, 3)].T"):
        for mat in ("vdw_R", "vdw_e", "q1q2"):
            old = f"{mat}[{idx}] = 0.0"
            new = f"{mat}[tuple({idx})] = 0.0"
            n += src.count(old)
            src = src.replace(old, new)
    if n != 9:
        raise RuntimeError(f"expected 9 exclusion-index sites in the reference, patched {n}")
    return src


def _accumulator_patch(src: str) -> str:
    old = "torch.tensor(0.0).float().to(self.device)"
    new = "torch.zeros((), dtype=x.dtype, device=self.device)"
    if src.count(old) != 6:
        raise RuntimeError(f"expected 6 accumulator sites, found {src.count(old)}")
    return src.replace(old, new)


def load(pair_exclusions: bool = True):
    """Return (utils_module, energy_module) of the reference, shims applied."""
    if not available():
        raise FileNotFoundError(f"molearn reference not found under {REF}")
    import pkg_resources  # noqa: deprecated but importable (setuptools 80)

    if not hasattr(np, "bool"):
        np.bool = bool  # shim 1
    pkg_resources.resource_filename = lambda *_a, **_k: PARAM_DIR  # shim 2
    for name in ("molearn", "molearn.loss_functions"):  # shim 3
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = []
            sys.modules[name] = pkg
    utils = _exec_module(
        "molearn.loss_functions.torch_protein_energy_utils",
        os.path.join(_LOSS_DIR, "torch_protein_energy_utils.py"),
        _pair_index_patch if pair_exclusions else None,  # shim 4
    )
    energy = _exec_module(
        "molearn.loss_functions.torch_protein_energy",
        os.path.join(_LOSS_DIR, "torch_protein_energy.py"),
        _accumulator_patch,  # shim 6
    )
    return utils, energy


def build_reference(frame, atominfo, pair_exclusions: bool = True, quiet: bool = True, **kw):
    """Construct the reference ``TorchProteinEnergy`` on CPU and capture its topology locals.

    frame: torch [3, N]; atominfo: object ndarray [N, 3] (name, resname, resid) - a copy is
    passed because the reference mutates it in place (utils.py:420-487).
    Returns (energy_object, captured_locals_dict, mutated_atominfo).
    """
    import torch

    utils, energy = load(pair_exclusions)
    captured = {}

    def tracer(frm, event, arg):
        if frm.f_code.co_name != "get_convolutions":
            return None

        def local(frm2, event2, arg2):
            if event2 == "return":
                for k in _CAPTURE_NAMES:
                    if k in frm2.f_locals:
                        captured[k] = frm2.f_locals[k]
            return local

        return local

    info = np.array(atominfo, dtype=object).copy()
    sink = io.StringIO()
    ctx = contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()
    sys.settrace(tracer)
    try:
        with ctx:
            obj = energy.TorchProteinEnergy(frame, info, device=torch.device("cpu"),
                                            method="roll", **kw)
    finally:
        sys.settrace(None)
    return obj, captured, info


def to_double(obj):
    """Cast every tensor attribute of a reference TorchProteinEnergy to float64."""
    import torch

    for k, v in list(vars(obj).items()):
        if torch.is_tensor(v) and v.is_floating_point():
            setattr(obj, k, v.double())
        elif isinstance(v, list) and v and torch.is_tensor(v[0]) and v[0].is_floating_point():
            setattr(obj, k, [t.double() for t in v])
    return obj
