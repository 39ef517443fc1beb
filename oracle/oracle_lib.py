"""ctypes wrapper of oracle/liboracle_mhd.so (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module.  The product package (pluto_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle_mhd.so")

RECON = {"plm": 0, "ppm": 1}
SOLVER = {"hlld": 0, "hll": 1, "roe": 2, "hllc": 3, "tvdlf": 4}
BC = {"periodic": 0, "outflow": 1, "reflective": 2, "eqtsymmetric": 3}
LIMITER = {"default": 0, "fl": 1, "mm": 2, "va": 3, "os": 4, "um": 5, "vl": 6, "mc": 7}
EMF = {"uct_contact": 0, "arith": 1, "uct0": 2, "uct_hll": 3}


class OracleConfig(C.Structure):
    _fields_ = [("dims", C.c_int), ("n", C.c_int * 3), ("recon", C.c_int),
                ("solver", C.c_int), ("rk_order", C.c_int), ("bc", C.c_int * 6),
                ("gamma", C.c_double), ("dx", C.c_double * 3),
                ("small_dn", C.c_double), ("small_pr", C.c_double),
                ("limiter", C.c_int), ("emf_average", C.c_int), ("shock_flattening", C.c_int),
                ("ctu", C.c_int), ("en_correction", C.c_int), ("body_force", C.c_int), ("grav", C.c_double * 3),
                ("char_limiting", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle_mhd.so"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(OracleConfig)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_body_force.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_set_grid.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_set_plm_coeffs.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.oracle_set_ppm_coeffs.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.oracle_set_body_potential.argtypes = [C.c_void_p] * 5
        L.oracle_nghost.argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        L.oracle_set_interior.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.oracle_get_interior.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.oracle_advance.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.oracle_advance.restype = C.c_int
        L.oracle_next_dt.argtypes = [C.c_double] * 4
        L.oracle_next_dt.restype = C.c_double
        L.oracle_debug_stop_after.argtypes = [C.c_void_p, C.c_int]
        L.oracle_tap.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_tap.restype = dp
        ip = C.POINTER(C.c_int)
        L.oracle_tap_shape.argtypes = [C.c_void_p, ip, ip, ip, ip]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


VC_NAMES_3D = ["rho", "vx1", "vx2", "vx3", "Bx1", "Bx2", "Bx3", "prs"]


class Oracle:
    """State container + stepper.  Arrays use the .dbl interior layout."""

    def __init__(self, dims, n, dx, recon="plm", solver="hlld", rk_order=2,
                 bc=("periodic",) * 6, gamma=5.0 / 3.0, limiter="default", emf="uct_contact", flatten=False, ctu=False, en_corr=False, grav=None,
                 char_lim=False):
        c = OracleConfig()
        c.body_force = 0 if grav is None else 1
        for d in range(3):
            c.grav[d] = 0.0 if grav is None else float(grav[d])
        c.ctu = 2 if ctu == "chtr" else (1 if ctu else 0)      # "chtr": TIME_STEPPING CHARACTERISTIC_TRACING
        c.en_correction = 1 if en_corr else 0
        c.char_limiting = 1 if char_lim else 0
        c.shock_flattening = 1 if flatten else 0
        c.limiter = LIMITER[limiter]
        c.emf_average = EMF[emf]
        c.dims = dims
        n = list(n) + [1] * (3 - len(n))
        if dims == 2:
            n[2] = 1
        for d in range(3):
            c.n[d] = n[d]
            c.dx[d] = dx[d] if d < len(dx) else 1.0
        c.recon = RECON[recon]
        c.solver = SOLVER[solver]
        c.rk_order = rk_order
        for s in range(6):
            c.bc[s] = BC[bc[s]]
        c.gamma = gamma
        c.small_dn = 1e-12
        c.small_pr = 1e-12
        self.cfg = c
        self.dims = dims
        self.n = tuple(n)
        self._h = lib().oracle_create(C.byref(c))
        self.ng = lib().oracle_nghost(self._h)

    def __del__(self):
        try:
            if self._h:
                lib().oracle_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_grid(self, dx1, dx2, dx3=None):
        """Non-uniform Cartesian grid: the zone widths grid->dx[d] of every direction, ghost zones included (T_d entries)."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (dx1, dx2, dx3)]
        lib().oracle_set_grid(self._h, *[a.ctypes.data if a is not None else None for a in arrs])

    def set_plm_coeffs(self, coeffs):
        """UNIFORM_CARTESIAN_GRID NO: per direction the six arrays (cp, cm, wp, wm, dp, dm) of PLM_CoefficientsGet, T entries each."""
        for d, six in enumerate(coeffs):
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in six]
            lib().oracle_set_plm_coeffs(self._h, d, *[a.ctypes.data for a in arrs])

    def set_ppm_coeffs(self, coeffs):
        """PARABOLIC on a non-uniform grid: per direction the four interface-weight arrays wp[i][-1 .. 2] of PPM_CoefficientsGet."""
        for d, four in enumerate(coeffs):
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in four]
            lib().oracle_set_ppm_coeffs(self._h, d, *[a.ctypes.data for a in arrs])

    def set_body_force(self, g1, g2, g3=None):
        """Static per-zone force: arrays [T3][T2][T1] (ghost zones included) of every component."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (g1, g2, g3)]
        self._gkeep = arrs
        lib().oracle_set_body_force(self._h, *[a.ctypes.data if a is not None else None for a in arrs])

    def set_body_potential(self, phic, pf1, pf2, pf3=None):
        """BODY_FORCE POTENTIAL: potential at the zone centres [T3][T2][T1] and at the faces (staggered Data layouts)."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (phic, pf1, pf2, pf3)]
        self._pkeep = arrs
        lib().oracle_set_body_potential(self._h, *[a.ctypes.data if a is not None else None for a in arrs])

    # ---- state I/O in "dump dict" form (names as in dbl.out) ----
    def set_state(self, dump: dict):
        n1, n2, n3 = self.n
        vc = np.zeros((8, n3, n2, n1))
        for iv, nm in enumerate(VC_NAMES_3D):
            if nm in dump:
                vc[iv] = dump[nm]
        b1 = np.ascontiguousarray(dump["Bx1s"], dtype=np.float64)
        b2 = np.ascontiguousarray(dump["Bx2s"], dtype=np.float64)
        b3 = np.ascontiguousarray(dump["Bx3s"], dtype=np.float64) if self.dims == 3 else None
        self._keep = (vc, b1, b2, b3)
        lib().oracle_set_interior(self._h, _dp(vc), _dp(b1), _dp(b2), _dp(b3))

    def get_state(self) -> dict:
        n1, n2, n3 = self.n
        vc = np.zeros((8, n3, n2, n1))
        b1 = np.zeros((n3, n2, n1 + 1))
        b2 = np.zeros((n3, n2 + 1, n1))
        b3 = np.zeros((n3 + 1, n2, n1)) if self.dims == 3 else None
        lib().oracle_get_interior(self._h, _dp(vc), _dp(b1), _dp(b2), _dp(b3))
        out = {}
        for iv, nm in enumerate(VC_NAMES_3D):
            if self.dims == 2 and nm in ("vx3", "Bx3"):
                continue
            out[nm] = vc[iv].copy()
        out["Bx1s"] = b1
        out["Bx2s"] = b2
        if self.dims == 3:
            out["Bx3s"] = b3
        return out

    def advance(self, dt: float):
        inv = C.c_double(0.0)
        mach = C.c_double(0.0)
        nfloor = lib().oracle_advance(self._h, dt, C.byref(inv), C.byref(mach))
        return inv.value, mach.value, nfloor

    def debug_stop_after(self, stage: int):
        lib().oracle_debug_stop_after(self._h, stage)

    def tap(self, name: str) -> np.ndarray:
        """Padded internal array [k+1][j+1][i+1] (copy)."""
        s3, s2, s1, tot = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib().oracle_tap_shape(self._h, C.byref(s3), C.byref(s2), C.byref(s1), C.byref(tot))
        p = lib().oracle_tap(self._h, name.encode())
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(s3.value, s2.value, s1.value)).copy()


def next_dt(inv_dt_hyp, cfl, cfl_max_var, dt):
    return lib().oracle_next_dt(inv_dt_hyp, cfl, cfl_max_var, dt)
