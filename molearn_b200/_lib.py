"""ctypes binding of ``libmolearn_b200.so`` (the C ABI declared in ``include/molearn_b200.h``).

There is no fallback: if the shared library is missing or cannot be loaded, importing any
compute entry point raises.  Build it with ``python -m molearn_b200.build``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p

from . import build as _build

_LIB = None

# every symbol include/molearn_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mlp_last_error": (c_char_p, []),
    "mlp_abi_version": (c_int, []),
    "mlp_topology_build": (c_int, [c_char_p, c_int32, POINTER(c_char_p), POINTER(c_char_p), c_void_p, c_uint32, POINTER(c_void_p)]),
    "mlp_topology_counts": (c_int, [c_void_p, c_void_p]),
    "mlp_topology_export": (c_int, [c_void_p] + [c_void_p] * 7),
    "mlp_topology_atoms": (c_int, [c_void_p] + [c_void_p] * 6),
    "mlp_topology_pairs14": (c_int, [c_void_p, c_void_p]),
    "mlp_topology_params14": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "mlp_topology_destroy": (None, [c_void_p]),
    "mlp_bonded_plan": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "mlp_nb_pair_table": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mlp_device_count": (c_int, []),
    "mlp_energy_create": (c_int, [c_void_p, c_int, c_uint32, POINTER(c_void_p)]),
    "mlp_energy_destroy": (None, [c_void_p]),
    "mlp_energy_workspace_bytes": (c_int64, [c_void_p, c_int64, c_uint32, c_int]),
    "mlp_energy_info": (c_int, [c_void_p, c_void_p]),
    "mlp_energy_nb_table_classes": (c_int, [c_void_p]),
    "mlp_energy_eval": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                c_void_p, c_int64, c_int64, c_int64, c_void_p, c_uint32, c_uint32,
                                c_void_p, c_int64, c_void_p]),
    "mlp_energy_last_launches": (c_int, [c_void_p]),
    "mlp_energy_profile": (c_int, [c_void_p, c_int]),
    "mlp_energy_profile_read": (c_int, [c_void_p, c_void_p, c_void_p]),
    "mlp_geometry_eval": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_int64,
                                  c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "mlp_internal_coords_eval": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_int64, c_void_p, c_int64,
                                         c_void_p, c_int64, c_void_p, c_void_p]),
    "mlp_probe_fp32": (c_int64, [c_void_p, c_int64, c_int, c_int, c_void_p]),
    "mlp_rmsd_eval": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_float,
                              c_void_p, c_int64, c_void_p, c_uint32, c_void_p]),
}

E_ARG, E_IO, E_KEY, E_PARSE, E_CUDA, E_NOGPU, E_WORKSPACE = -1, -2, -3, -4, -5, -6, -7
TERM_BOND, TERM_ANGLE, TERM_TORSION, TERM_NB, TERM_BONDED, TERM_ALL = 1, 2, 4, 8, 7, 15
NB_REPULSIVE, NB_FULL = 0, 1
TOPO_FIX_H = 1
EVAL_ACCUMULATE_GRAD = 1
RMSD_ALIGN = 1
RMSD_PAIRED = 2


def library_path() -> str:
    """The in-tree library; ``MOLEARN_B200_LIB`` points at another build of the same ABI (A/B timing of kernel variants)."""
    return os.environ.get("MOLEARN_B200_LIB") or _build.LIB


def lib():
    """Load (once) and return the shared library with argtypes set."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            try:                                   # a fresh checkout: compile once (nvcc, ~20 s); never a CPU path
                _build.build_library()
            except Exception as exc:
                raise ImportError(
                    f"{path} is missing and could not be built ({exc}): the molearn_b200 CUDA extension is required, "
                    "there is no CPU fallback.  Run `python -m molearn_b200.build` (needs nvcc).") from exc
        L = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        if L.mlp_abi_version() != 1:
            raise ImportError(f"{path}: ABI version {L.mlp_abi_version()} != 1, rebuild the extension")
        _LIB = L
    return _LIB


def check(rc: int):
    """Turn a negative return code into the exception the reference would have raised."""
    if rc >= 0:
        return rc
    msg = lib().mlp_last_error().decode(errors="replace")
    if rc == E_KEY:
        raise KeyError(msg)
    if rc == E_IO:
        raise FileNotFoundError(msg)
    if rc == E_ARG:
        raise ValueError(msg)
    raise RuntimeError(f"molearn_b200 error {rc}: {msg}")
