"""Host-side mirror of the reference's step interface.

``GpuStepper.advance(dt)`` is AdvanceStep (reference Src/Time_Stepping/
rk_step.c:27) on device-resident state; ``Integrator`` is the time loop of
main() (Src/main.c:133-243): clip dt to tstop, Integrate, g_time += g_dt,
g_dt = NextTimeStep.  State dictionaries use the reference's dbl.out names
(rho vx1 vx2 [vx3] Bx1 Bx2 [Bx3] prs Bx1s Bx2s [Bx3s]) and interior shapes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

VC_NAMES = ["rho", "vx1", "vx2", "vx3", "Bx1", "Bx2", "Bx3", "prs"]


@dataclass
class StepInfo:
    inv_dt_hyp: float
    max_mach: float
    floor_events: int
    nan_events: int


class PlutoGpuError(RuntimeError):
    pass


class GpuStepper:
    """One block of the domain resident on one GPU."""

    def __init__(self, dims, n, dx, recon="plm", solver="hlld", rk_order=2,
                 bc=("periodic",) * 6, gamma=5.0 / 3.0, arith="exact", device=0,
                 small_dn=1e-12, small_pr=1e-12, lib_path=None, limiter="default", emf="uct_contact",
                 flatten=False, ctu=False, en_corr=False, grav=None, potential=False, char_lim=False):
        self.L = _lib.load_library(lib_path)
        c = _lib.PlutoGpuConfig()
        n = list(n) + [1] * (3 - len(n))
        if dims == 2:
            n[2] = 1
        c.dims = dims
        for d in range(3):
            c.n[d] = int(n[d])
            c.dx[d] = float(dx[d]) if d < len(dx) else 1.0
        c.recon = _lib.RECON[recon]
        c.solver = _lib.SOLVER[solver]
        c.rk_order = rk_order
        for s in range(6):
            c.bc[s] = _lib.BC[bc[s]]
        c.arith = _lib.ARITH[arith]
        c.device = device
        c.gamma = gamma
        c.small_dn = small_dn
        c.small_pr = small_pr
        c.limiter = _lib.LIMITER[limiter]           # LIMITER (plm only): default | fl mm va os um vl mc
        c.emf_average = _lib.EMF[emf]               # CT_EMF_AVERAGE: uct_contact | arith | uct0 | uct_hll
        c.shock_flattening = 1 if flatten else 0    # SHOCK_FLATTENING MULTID (plm only)
        c.time_stepping = 2 if ctu == "chtr" else (1 if ctu else 0)   # TIME_STEPPING: RK2/RK3 (rk_order) | HANCOCK | CHARACTERISTIC_TRACING (ctu="chtr")
        c.en_correction = 1 if en_corr else 0       # CT_EN_CORRECTION YES
        c.char_limiting = 1 if char_lim else 0      # CHAR_LIMITING YES (2-D, plm, RK)
        # BODY_FORCE: VECTOR (bit 0) with the uniform acceleration grav, POTENTIAL (bit 1, set_body_potential)
        c.body_force = (0 if grav is None else 1) | (2 if potential else 0)
        for d in range(3):
            c.grav[d] = 0.0 if grav is None else float(grav[d])
        self.cfg = c
        self.dims = dims
        self.n = tuple(n)
        self.rk_order = rk_order
        self._h = C.c_void_p()
        if self.L.pluto_gpu_create(C.byref(c), C.byref(self._h)) != 0:
            self._h = None
            raise PlutoGpuError(_lib.last_error(self.L))
        self.ng = self.L.pluto_gpu_nghost(self._h)
        self.nstages = self.L.pluto_gpu_nstages(self._h)    # Boundary calls (halo exchanges) per step

    def close(self):
        if getattr(self, "_h", None):
            self.L.pluto_gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PlutoGpuError(_lib.last_error(self.L))

    def set_grid(self, dx1, dx2, dx3=None):
        """Non-uniform Cartesian grid: the zone widths grid->dx[d] of every direction, ghost zones included (n[d] + 2 nghost
        entries).  RK2 / RK3 with LINEAR reconstruction; call before the first step."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (dx1, dx2, dx3)]
        for d, a in enumerate(arrs[:self.dims]):
            if a is None or a.size != self.n[d] + 2 * self.ng:
                raise ValueError(f"set_grid: dx{d+1} needs {self.n[d] + 2 * self.ng} entries")
        self._check(self.L.pluto_gpu_set_grid(self._h, *[a.ctypes.data if a is not None else None for a in arrs]))

    def set_plm_coeffs(self, coeffs):
        """UNIFORM_CARTESIAN_GRID NO: per direction the six arrays (cp, cm, wp, wm, dp, dm) PLM_CoefficientsGet returns
        (n[d] + 2 nghost entries each)."""
        for d, six in enumerate(coeffs):
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in six]
            if len(arrs) != 6 or any(a.size != self.n[d] + 2 * self.ng for a in arrs):
                raise ValueError(f"set_plm_coeffs: direction {d+1} needs six arrays of {self.n[d] + 2 * self.ng} entries")
            self._check(self.L.pluto_gpu_set_plm_coeffs(self._h, d, *[a.ctypes.data for a in arrs]))

    def set_ppm_coeffs(self, coeffs):
        """PARABOLIC on a non-uniform grid: per direction the four interface-weight arrays wp[i][-1 .. 2] PPM_CoefficientsGet returns
        (n[d] + 2 nghost entries each)."""
        for d, four in enumerate(coeffs):
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in four]
            if len(arrs) != 4 or any(a.size != self.n[d] + 2 * self.ng for a in arrs):
                raise ValueError(f"set_ppm_coeffs: direction {d+1} needs four arrays of {self.n[d] + 2 * self.ng} entries")
            self._check(self.L.pluto_gpu_set_ppm_coeffs(self._h, d, *[a.ctypes.data for a in arrs]))

    def set_body_force(self, g1, g2, g3=None):
        """Static position-dependent force (BodyForceVector at the zone centres): arrays [T3][T2][T1] incl. ghost zones.
        The stepper must have been created with grav=... (BODY_FORCE VECTOR)."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (g1, g2, g3)]
        self._check(self.L.pluto_gpu_set_body_force(self._h, *[a.ctypes.data if a is not None else None for a in arrs]))

    def set_body_potential(self, phic, pf1, pf2, pf3=None):
        """BODY_FORCE POTENTIAL: BodyForcePotential at the zone centres [T3][T2][T1] and at the faces of every direction
        (staggered Data layouts, one more face starting at -1/2).  Create the stepper with potential=True."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (phic, pf1, pf2, pf3)]
        self._check(self.L.pluto_gpu_set_body_potential(self._h, *[a.ctypes.data if a is not None else None for a in arrs]))

    # ---- interior (.dbl) layout -------------------------------------------
    def interior_buffers(self, pinned=False):
        """Host arrays in the interior layout: (vc[8,n3,n2,n1], bx1s, bx2s, bx3s|None)."""
        n1, n2, n3 = self.n
        shapes = [(8, n3, n2, n1), (n3, n2, n1 + 1), (n3, n2 + 1, n1)]
        if self.dims == 3:
            shapes.append((n3 + 1, n2, n1))
        if pinned:
            import torch
            bufs = [torch.zeros(s, dtype=torch.float64).pin_memory().numpy() for s in shapes]
        else:
            bufs = [np.zeros(s) for s in shapes]
        if self.dims == 2:
            bufs.append(None)
        return bufs

    def upload_interior(self, vc, b1, b2, b3=None):
        p = lambda a: a.ctypes.data if a is not None else None
        for a in (vc, b1, b2, b3):
            assert a is None or (a.dtype == np.float64 and a.flags["C_CONTIGUOUS"])
        self._check(self.L.pluto_gpu_upload_interior(self._h, p(vc), p(b1), p(b2), p(b3)))

    def download_interior(self, vc, b1, b2, b3=None):
        p = lambda a: a.ctypes.data if a is not None else None
        self._check(self.L.pluto_gpu_download_interior(self._h, p(vc), p(b1), p(b2), p(b3)))

    def set_state(self, dump: dict):
        vc, b1, b2, b3 = self.interior_buffers()
        for iv, nm in enumerate(VC_NAMES):
            if nm in dump:
                vc[iv] = dump[nm]
        b1[...] = dump["Bx1s"]
        b2[...] = dump["Bx2s"]
        if self.dims == 3:
            b3[...] = dump["Bx3s"]
        self.upload_interior(vc, b1, b2, b3)

    def get_state(self) -> dict:
        vc, b1, b2, b3 = self.interior_buffers()
        self.download_interior(vc, b1, b2, b3)
        out = {}
        for iv, nm in enumerate(VC_NAMES):
            if self.dims == 2 and nm in ("vx3", "Bx3"):
                continue
            out[nm] = vc[iv].copy()
        out["Bx1s"], out["Bx2s"] = b1, b2
        if self.dims == 3:
            out["Bx3s"] = b3
        return out

    # ---- reference Data layout (with ghost zones) --------------------------
    def data_buffers(self, pinned=False):
        n1, n2, n3 = self.n
        g = self.ng
        T1, T2 = n1 + 2 * g, n2 + 2 * g
        T3 = n3 + 2 * g if self.dims == 3 else 1
        nvar = 8 if self.dims == 3 else 6
        shapes = [(nvar, T3, T2, T1), (T3, T2, T1 + 1), (T3, T2 + 1, T1)]
        if self.dims == 3:
            shapes.append((T3 + 1, T2, T1))
        if pinned:
            import torch
            bufs = [torch.zeros(s, dtype=torch.float64).pin_memory().numpy() for s in shapes]
        else:
            bufs = [np.zeros(s) for s in shapes]
        if self.dims == 2:
            bufs.append(None)
        return bufs

    def upload_data(self, Vc, s1, s2, s3=None):
        p = lambda a: a.ctypes.data if a is not None else None
        self._check(self.L.pluto_gpu_upload_data(self._h, p(Vc), p(s1), p(s2), p(s3)))

    def download_data(self, Vc, s1, s2, s3=None):
        p = lambda a: a.ctypes.data if a is not None else None
        self._check(self.L.pluto_gpu_download_data(self._h, p(Vc), p(s1), p(s2), p(s3)))

    def advance_data(self, dt, Vc, s1, s2, s3=None) -> StepInfo:
        """AdvanceStep on HOST Data arrays: upload, step, download."""
        p = lambda a: a.ctypes.data if a is not None else None
        info = _lib.PlutoGpuStepInfo()
        self._check(self.L.pluto_gpu_advance_data(self._h, dt, p(Vc), p(s1), p(s2), p(s3), C.byref(info)))
        return StepInfo(info.inv_dt_hyp, info.max_mach, info.floor_events, info.nan_events)

    # ---- the step -----------------------------------------------------------
    def advance(self, dt: float) -> StepInfo:
        info = _lib.PlutoGpuStepInfo()
        self._check(self.L.pluto_gpu_advance(self._h, dt, C.byref(info)))
        return StepInfo(info.inv_dt_hyp, info.max_mach, info.floor_events, info.nan_events)

    # ---- NextTimeStep on the device: steps enqueued back to back ------------------
    def set_dt(self, dt: float):
        self._check(self.L.pluto_gpu_set_dt(self._h, dt))

    def advance_async(self, cfl: float, cfl_max_var: float = 1.1):
        """Enqueue one step with the device's dt; the next dt is computed on the device."""
        self._check(self.L.pluto_gpu_advance_async(self._h, cfl, cfl_max_var))

    def next_dt_async(self, cfl: float, cfl_max_var: float = 1.1):
        self._check(self.L.pluto_gpu_next_dt_async(self._h, cfl, cfl_max_var))

    def reduction_slots(self) -> int:
        p = C.c_void_p()
        self._check(self.L.pluto_gpu_reduction_slots(self._h, C.byref(p)))
        return int(p.value)

    def sync_results(self, max_steps: int = 4096):
        """Wait for the enqueued steps: ([dt used], [StepInfo], dt of the next step)."""
        infos = (_lib.PlutoGpuStepInfo * max_steps)()
        dts = (C.c_double * max_steps)()
        n, dtn = C.c_int(0), C.c_double(0.0)
        self._check(self.L.pluto_gpu_sync_results(self._h, max_steps, infos, dts, C.byref(n), C.byref(dtn)))
        out = [StepInfo(infos[q].inv_dt_hyp, infos[q].max_mach, infos[q].floor_events, infos[q].nan_events)
               for q in range(n.value)]
        return [dts[q] for q in range(n.value)], out, dtn.value

    # ---- output / restart / diagnostics from the device state ---------------------
    def write_dbl(self, directory: str, nfile: int, t: float, dt: float, nstep: int):
        """data.NNNN.dbl + dbl.out line in the reference's single_file format (Src/write_data.c)."""
        self._check(self.L.pluto_gpu_write_dbl(self._h, directory.encode(), nfile, t, dt, nstep))

    def write_flt(self, directory: str, nfile: int, t: float, dt: float, nstep: int):
        """data.NNNN.flt + flt.out: the cell-centred variables in single precision (Src/write_data.c:178-206)."""
        self._check(self.L.pluto_gpu_write_flt(self._h, directory.encode(), nfile, t, dt, nstep))

    def write_vtk(self, directory: str, nfile: int, t: float, dt: float, nstep: int, xl):
        """data.NNNN.vtk + vtk.out (Src/write_vtk.c); xl = node coordinates per direction (n+1 values each)."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (list(xl) + [None] * 3)[:3]]
        self._check(self.L.pluto_gpu_write_vtk(self._h, directory.encode(), nfile, t, dt, nstep,
                                               *[a.ctypes.data if a is not None else None for a in arrs]))

    def read_dbl(self, path: str):
        self._check(self.L.pluto_gpu_read_dbl(self._h, path.encode()))

    def analysis(self) -> dict:
        out = (C.c_double * 8)()
        self._check(self.L.pluto_gpu_analysis(self._h, out))
        return dict(mass=out[0], e_kin=out[1], e_mag=out[2], e_th=out[3],
                    mom=(out[4], out[5], out[6]), max_divb=out[7])

    def boundary(self):
        self._check(self.L.pluto_gpu_boundary(self._h))

    def step_begin(self):
        self._check(self.L.pluto_gpu_step_begin(self._h))

    def stage(self, stage, dt):
        self._check(self.L.pluto_gpu_stage(self._h, stage, dt))

    def stage_shell(self, stage, dt):
        self._check(self.L.pluto_gpu_stage_shell(self._h, stage, dt))

    def stage_interior(self, stage):
        self._check(self.L.pluto_gpu_stage_interior(self._h, stage))

    def boundary_dim(self, stage, dim):
        self._check(self.L.pluto_gpu_boundary_dim(self._h, stage, dim))

    def step_end(self) -> StepInfo:
        info = _lib.PlutoGpuStepInfo()
        self._check(self.L.pluto_gpu_step_end(self._h, C.byref(info)))
        return StepInfo(info.inv_dt_hyp, info.max_mach, info.floor_events, info.nan_events)

    def halo_doubles(self, dim) -> int:
        return int(self.L.pluto_gpu_halo_doubles(self._h, dim))

    def halo_pack(self, stage, dim, send_lo_ptr, send_hi_ptr):
        self._check(self.L.pluto_gpu_halo_pack(self._h, stage, dim, send_lo_ptr, send_hi_ptr))

    def halo_unpack(self, stage, dim, recv_lo_ptr, recv_hi_ptr):
        self._check(self.L.pluto_gpu_halo_unpack(self._h, stage, dim, recv_lo_ptr, recv_hi_ptr))

    def read_field(self, name: str) -> np.ndarray:
        """Debug tap: the whole padded device array [k+off][j+1][i+1] of a field."""
        dev = C.POINTER(C.c_double)()
        shape = (C.c_longlong * 3)()
        off = (C.c_int * 3)()
        self._check(self.L.pluto_gpu_field(self._h, name.encode(), C.byref(dev), C.byref(shape), C.byref(off)))
        out = np.zeros((shape[2], shape[1], shape[0]))
        self._check(self.L.pluto_gpu_read_field(self._h, name.encode(), out.ctypes.data))
        return out

    # ---- all-neighbour halo plan --------------------------------------------
    def halo_nbr_doubles(self, off) -> int:
        o = (C.c_int * 3)(*off)
        return int(self.L.pluto_gpu_halo_nbr_doubles(self._h, C.byref(o)))

    def halo_plan(self, offsets, send_ptrs, recv_ptrs):
        n = len(offsets)
        flat = (C.c_int * (3 * n))(*[c for o in offsets for c in o])
        sp = (C.c_void_p * n)(*send_ptrs)
        rp = (C.c_void_p * n)(*recv_ptrs)
        self._check(self.L.pluto_gpu_halo_plan(self._h, n, flat, sp, rp))

    def halo_plan_stage(self, stage, offsets, send_ptrs, recv_ptrs):
        n = len(offsets)
        flat = (C.c_int * (3 * n))(*[c for o in offsets for c in o])
        sp = (C.c_void_p * n)(*send_ptrs)
        rp = (C.c_void_p * n)(*recv_ptrs)
        self._check(self.L.pluto_gpu_halo_plan_stage(self._h, stage, n, flat, sp, rp))

    def halo_signal(self, stream_ptr, peer_counter_ptrs, value):
        n = len(peer_counter_ptrs)
        pp = (C.c_void_p * max(n, 1))(*peer_counter_ptrs)
        self._check(self.L.pluto_gpu_halo_signal(self._h, stream_ptr, n, pp, value))

    def halo_wait(self, stream_ptr, n, counters_ptr, value):
        self._check(self.L.pluto_gpu_halo_wait(self._h, stream_ptr, n, counters_ptr, value))

    def halo_pack_all(self, stage):
        self._check(self.L.pluto_gpu_halo_pack_all(self._h, stage))

    def halo_pack_all_on(self, stage, stream_ptr):
        self._check(self.L.pluto_gpu_halo_pack_all_on(self._h, stage, stream_ptr))

    def halo_unpack_all(self, stage):
        self._check(self.L.pluto_gpu_halo_unpack_all(self._h, stage))

    def timing(self, enable: bool):
        self.L.pluto_gpu_timing(self._h, 1 if enable else 0)

    def timing_report(self) -> dict:
        """{class name: (total ms, launches)} accumulated since timing(True)."""
        out = {}
        for c in range(8):
            nm, ms, cnt = C.c_char_p(), C.c_double(), C.c_longlong()
            if self.L.pluto_gpu_timing_get(self._h, c, C.byref(nm), C.byref(ms), C.byref(cnt)) == 0:
                out[nm.value.decode()] = (ms.value, cnt.value)
        return out

    @property
    def stream(self) -> int:
        return int(self.L.pluto_gpu_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.L.pluto_gpu_launch_count(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self.L.pluto_gpu_device_bytes(self._h))

    def next_dt(self, inv_dt_hyp, cfl, cfl_max_var, dt) -> float:
        return float(self.L.pluto_gpu_next_dt(inv_dt_hyp, cfl, cfl_max_var, dt))


class MultiGpuStepper:
    """AdvanceStep on a domain cut into blocks, one per GPU, driven from THIS thread through the C ABI alone
    (pluto_gpu_multi_*): what a single-threaded C host uses to reach the GPUs of a box without MPI.  Host arrays are the
    reference's Data arrays of the whole domain (ghost zones included)."""

    def __init__(self, dims, n, dx, grid, devices=None, recon="plm", solver="hlld", rk_order=2, bc=("periodic",) * 6,
                 gamma=5.0 / 3.0, arith="exact", lib_path=None, limiter="default", emf="uct_contact", flatten=False, ctu=False,
                 en_corr=False, char_lim=False):
        self.L = _lib.load_library(lib_path)
        c = _lib.PlutoGpuConfig()
        n = list(n) + [1] * (3 - len(n))
        if dims == 2:
            n[2] = 1
        c.dims = dims
        for d in range(3):
            c.n[d] = int(n[d])
            c.dx[d] = float(dx[d]) if d < len(dx) else 1.0
        c.recon, c.solver, c.rk_order = _lib.RECON[recon], _lib.SOLVER[solver], rk_order
        for s in range(6):
            c.bc[s] = _lib.BC[bc[s]]
        c.arith, c.gamma, c.small_dn, c.small_pr = _lib.ARITH[arith], gamma, 1e-12, 1e-12
        c.limiter, c.emf_average = _lib.LIMITER[limiter], _lib.EMF[emf]
        c.shock_flattening, c.time_stepping = (1 if flatten else 0), (2 if ctu == "chtr" else (1 if ctu else 0))
        c.en_correction, c.char_limiting = (1 if en_corr else 0), (1 if char_lim else 0)
        self.dims, self.n = dims, tuple(n)
        g = (C.c_int * 3)(*(list(grid) + [1] * (3 - len(grid))))
        nb = g[0] * g[1] * g[2]
        dev = (C.c_int * nb)(*(devices if devices is not None else [0] * nb))
        self._h = C.c_void_p()
        if self.L.pluto_gpu_multi_create(C.byref(c), C.byref(g), dev, C.byref(self._h)) != 0:
            self._h = None
            raise PlutoGpuError(_lib.last_error(self.L))
        self.ng = self.L.pluto_gpu_multi_nghost(self._h)
        self.nblocks = self.L.pluto_gpu_multi_nblocks(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self.L.pluto_gpu_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PlutoGpuError(_lib.last_error(self.L))

    def set_grid(self, dx1, dx2, dx3=None):
        """Zone widths of the WHOLE domain per direction (ghost zones included); every block takes its slice."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (dx1, dx2, dx3)]
        self._check(self.L.pluto_gpu_multi_set_grid(self._h, *[a.ctypes.data if a is not None else None for a in arrs]))

    def set_ppm_coeffs(self, coeffs):
        """PARABOLIC on a non-uniform grid: per direction the four interface-weight arrays of the WHOLE domain."""
        for d, four in enumerate(coeffs):
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in four]
            self._check(self.L.pluto_gpu_multi_set_ppm_coeffs(self._h, d, *[a.ctypes.data for a in arrs]))

    def set_body_force(self, g1, g2, g3=None):
        """BodyForceVector at the zone centres of the WHOLE domain ([T3][T2][T1], ghost zones included); every block takes its piece."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (g1, g2, g3)]
        self._check(self.L.pluto_gpu_multi_set_body_force(self._h, *[a.ctypes.data if a is not None else None for a in arrs]))

    def set_body_potential(self, phic, pf1, pf2, pf3=None):
        """BodyForcePotential of the WHOLE domain: zone centres and the faces of every direction (staggered Data layouts)."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (phic, pf1, pf2, pf3)]
        self._check(self.L.pluto_gpu_multi_set_body_potential(self._h, *[a.ctypes.data if a is not None else None for a in arrs]))

    def set_plm_coeffs(self, coeffs):
        """UNIFORM_CARTESIAN_GRID NO: per direction the six weight arrays of the WHOLE domain."""
        for d, six in enumerate(coeffs):
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in six]
            self._check(self.L.pluto_gpu_multi_set_plm_coeffs(self._h, d, *[a.ctypes.data for a in arrs]))

    def data_buffers(self):
        n1, n2, n3 = self.n
        g = self.ng
        T1, T2 = n1 + 2 * g, n2 + 2 * g
        T3 = n3 + 2 * g if self.dims == 3 else 1
        shapes = [(8 if self.dims == 3 else 6, T3, T2, T1), (T3, T2, T1 + 1), (T3, T2 + 1, T1)]
        if self.dims == 3:
            shapes.append((T3 + 1, T2, T1))
        bufs = [np.zeros(s) for s in shapes]
        if self.dims == 2:
            bufs.append(None)
        return bufs

    def upload_data(self, Vc, s1, s2, s3=None):
        p = lambda a: a.ctypes.data if a is not None else None
        self._check(self.L.pluto_gpu_multi_upload_data(self._h, p(Vc), p(s1), p(s2), p(s3)))

    def download_data(self, Vc, s1, s2, s3=None):
        p = lambda a: a.ctypes.data if a is not None else None
        self._check(self.L.pluto_gpu_multi_download_data(self._h, p(Vc), p(s1), p(s2), p(s3)))

    def advance(self, dt: float) -> StepInfo:
        info = _lib.PlutoGpuStepInfo()
        self._check(self.L.pluto_gpu_multi_advance(self._h, dt, C.byref(info)))
        return StepInfo(info.inv_dt_hyp, info.max_mach, info.floor_events, info.nan_events)

    def set_state(self, dump: dict):
        """Interior (.dbl) state -> the Data arrays of the whole domain (ghost zones left to the exchange / Boundary)."""
        Vc, s1, s2, s3 = self.data_buffers()
        g = self.ng
        names = VC_NAMES if self.dims == 3 else ["rho", "vx1", "vx2", "Bx1", "Bx2", "prs"]
        k = slice(g, -g) if self.dims == 3 else slice(None)
        for iv, nm in enumerate(names):
            Vc[iv][k, g:-g, g:-g] = dump[nm]
        s1[k, g:-g, g:-g] = dump["Bx1s"]
        s2[k, g:-g, g:-g] = dump["Bx2s"]
        if self.dims == 3:
            s3[g:-g, g:-g, g:-g] = dump["Bx3s"]
        self.upload_data(Vc, s1, s2, s3)

    def get_state(self) -> dict:
        Vc, s1, s2, s3 = self.data_buffers()
        self.download_data(Vc, s1, s2, s3)
        g = self.ng
        names = VC_NAMES if self.dims == 3 else ["rho", "vx1", "vx2", "Bx1", "Bx2", "prs"]
        k = slice(g, -g) if self.dims == 3 else slice(None)
        out = {nm: Vc[iv][k, g:-g, g:-g].copy() for iv, nm in enumerate(names)}
        out["Bx1s"] = s1[k, g:-g, g:-g].copy()
        out["Bx2s"] = s2[k, g:-g, g:-g].copy()
        if self.dims == 3:
            out["Bx3s"] = s3[g:-g, g:-g, g:-g].copy()
        return out


class Integrator:
    """The reference's main loop around AdvanceStep (Src/main.c:133-243)."""

    def __init__(self, stepper: GpuStepper, cfl: float, cfl_max_var: float = 1.1,
                 first_dt: float = 1e-4, tstop: float = 1e30):
        self.s = stepper
        self.cfl, self.cfl_max_var = cfl, cfl_max_var
        self.dt = first_dt
        self.t = 0.0
        self.tstop = tstop
        self.nstep = 0
        self.max_mach = 0.0
        self.dt_history = [first_dt]

    def step(self) -> StepInfo:
        if self.t + self.dt >= self.tstop * (1.0 - 1e-8):      # main.c:145-148
            self.dt = self.tstop - self.t
        info = self.s.advance(self.dt)
        if info.nan_events:
            raise PlutoGpuError(f"step {self.nstep}: {info.nan_events} zones are not finite")
        self.t += self.dt                                       # main.c:215
        self.dt = self.s.next_dt(info.inv_dt_hyp, self.cfl, self.cfl_max_var, self.dt)  # main.c:236
        self.max_mach = info.max_mach
        self.nstep += 1
        self.dt_history.append(self.dt)
        return info

    def run(self, nsteps: int):
        for _ in range(nsteps):
            self.step()
        return self
