"""TEST INFRASTRUCTURE ONLY - ctypes front-end of ``oracle/energy_oracle.c`` (the fp64 CPU
restatement of the reference energy; see that file's header for the reference lines it follows).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg import this.
The product (``molearn_b200``) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "liboracle.so")
SRC = os.path.join(HERE, "energy_oracle.c")

_lib = None


def build(force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        try:       # the O(N^2) non-bonded loop uses the host threads when OpenMP is there (51 432-atom parity cases)
            subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"], stderr=subprocess.DEVNULL)
        except (subprocess.CalledProcessError, OSError):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        L.oracle_nonbonded.restype = ctypes.c_double
        L.oracle_rmsd.restype = ctypes.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def bonded(x, topo, want_grad=True, fp32_params=True):
    """x [N,3]; topo: dict with bond_idxs/angle_idxs/torsion_idxs + *_para (golden layout).
    Returns (e[3], g[3,N,3] or None) in fp64.  fp32_params: round the parameters to fp32 first,
    as the reference does when it moves them to the device (torch_protein_energy.py:87-88,
    105-106, 123 ``.float()``) - its fp64 runs keep those rounded values."""
    x = _c(x, np.float64)
    n = x.shape[0]
    b, a, t = (_c(topo[k], np.int32) for k in ("bond_idxs", "angle_idxs", "torsion_idxs"))
    bp, ap, tp = (topo[k] for k in ("bond_para", "angle_para", "torsion_para"))
    if fp32_params:
        bp, ap, tp = (np.asarray(v, dtype=np.float32) for v in (bp, ap, tp))
    bp, ap, tp = (_c(v, np.float64) for v in (bp, ap, tp))
    e = np.zeros(3)
    g = np.zeros((3, n, 3)) if want_grad else None
    lib().oracle_bonded(ctypes.c_int32(n), _p(x), ctypes.c_int32(len(b)), _p(b), _p(bp),
                        ctypes.c_int32(len(a)), _p(a), _p(ap),
                        ctypes.c_int32(len(t)), _p(t), _p(tp), ctypes.c_int32(tp.shape[2] if tp.ndim == 3 else 0),
                        _p(e), _p(g))
    return e, g


def nonbonded(x, R, eps, q, excl, full=False, want_grad=True):
    """x [N,3]; R, eps, q [N] (fp32 reference values); excl [ne,2] sorted unique i<j."""
    x = _c(x, np.float64)
    n = x.shape[0]
    R, eps, q = (_c(np.asarray(v, dtype=np.float32), np.float64) for v in (R, eps, q))
    excl = _c(excl, np.int32).reshape(-1, 2)
    g = np.zeros((n, 3)) if want_grad else None
    e = lib().oracle_nonbonded(ctypes.c_int32(n), _p(x), _p(R), _p(eps), _p(q),
                               ctypes.c_int32(len(excl)), _p(excl), ctypes.c_int32(int(full)), _p(g))
    return e, g


def rmsd(a, b):
    a, b = _c(a, np.float64), _c(b, np.float64)
    return lib().oracle_rmsd(ctypes.c_int32(a.shape[0]), _p(a), _p(b))


def exclusions(topo):
    """Unique sorted (i<j) 1-2 / 1-3 / 1-4 pairs from golden-layout index lists."""
    p = np.concatenate([topo["bond_idxs"], topo["angle_idxs"][:, (0, 2)], topo["torsion_idxs"][:, (0, 3)]])
    p = np.sort(p, axis=1)
    p = p[p[:, 0] != p[:, 1]]
    return np.unique(p, axis=0).astype(np.int32)


def energy_all(x, topo, full=False, want_grad=True):
    """All four terms of one conformation: (e[4], g[4,N,3])."""
    e3, g3 = bonded(x, topo, want_grad)
    enb, gnb = nonbonded(x, topo["atom_R"], topo["atom_e"], topo["atom_charges"], exclusions(topo), full, want_grad)
    e = np.concatenate([e3, [enb]])
    g = np.concatenate([g3, gnb[None]]) if want_grad else None
    return e, g
