"""ctypes binding of include/pluto_gpu.h (the C ABI is the drop-in boundary)."""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libpluto_gpu.so")

RECON = {"plm": 0, "linear": 0, "ppm": 1, "parabolic": 1}
SOLVER = {"hlld": 0, "hll": 1, "roe": 2, "hllc": 3, "tvdlf": 4}
BC = {"periodic": 0, "outflow": 1, "reflective": 2, "shared": 3, "eqtsymmetric": 4}
LIMITER = {"default": 0, "fl": 1, "mm": 2, "va": 3, "os": 4, "um": 5, "vl": 6, "mc": 7}      # LIMITER
EMF = {"uct_contact": 0, "arith": 1, "uct0": 2, "uct_hll": 3}                                                 # CT_EMF_AVERAGE
ARITH = {"exact": 0, "fast": 1}
TIME_STEPPING = {"rk": 0, "hancock": 1, "chtr": 2}                                                                       # TIME_STEPPING

# every symbol include/pluto_gpu.h declares (checked by tests/test_cabi.py)
SYMBOLS = [
    "pluto_gpu_create", "pluto_gpu_destroy", "pluto_gpu_last_error", "pluto_gpu_nghost", "pluto_gpu_nstages", "pluto_gpu_set_body_force", "pluto_gpu_set_body_potential", "pluto_gpu_set_grid", "pluto_gpu_set_plm_coeffs", "pluto_gpu_set_ppm_coeffs", "pluto_gpu_multi_set_ppm_coeffs",
    "pluto_gpu_upload_interior", "pluto_gpu_download_interior", "pluto_gpu_upload_data",
    "pluto_gpu_download_data", "pluto_gpu_advance", "pluto_gpu_advance_data",
    "pluto_gpu_boundary", "pluto_gpu_next_dt", "pluto_gpu_halo_doubles", "pluto_gpu_halo_pack",
    "pluto_gpu_halo_unpack", "pluto_gpu_boundary_dim", "pluto_gpu_step_begin", "pluto_gpu_stage",
    "pluto_gpu_step_end", "pluto_gpu_stream", "pluto_gpu_launch_count", "pluto_gpu_device_bytes",
    "pluto_gpu_field", "pluto_gpu_read_field", "pluto_gpu_timing", "pluto_gpu_timing_get", "pluto_gpu_measure_fp64", "pluto_gpu_selftest_arith", "pluto_gpu_halo_nbr_doubles", "pluto_gpu_halo_plan",
    "pluto_gpu_halo_pack_all", "pluto_gpu_halo_unpack_all", "pluto_gpu_halo_pack_all_on", "pluto_gpu_halo_plan_stage",
    "pluto_gpu_device_count", "pluto_gpu_multi_create", "pluto_gpu_multi_destroy", "pluto_gpu_multi_nghost", "pluto_gpu_multi_nblocks", "pluto_gpu_multi_upload_data",
    "pluto_gpu_multi_download_data", "pluto_gpu_multi_advance", "pluto_gpu_multi_advance_data", "pluto_gpu_multi_set_grid", "pluto_gpu_multi_set_plm_coeffs",
    "pluto_gpu_multi_set_body_force", "pluto_gpu_multi_set_body_potential",
    "pluto_gpu_ipc_alloc", "pluto_gpu_ipc_open", "pluto_gpu_ipc_close", "pluto_gpu_ipc_free", "pluto_gpu_halo_signal", "pluto_gpu_halo_wait",
    "pluto_gpu_stage_shell", "pluto_gpu_stage_interior",
    "pluto_gpu_set_dt", "pluto_gpu_advance_async", "pluto_gpu_next_dt_async", "pluto_gpu_reduction_slots",
    "pluto_gpu_sync_results", "pluto_gpu_write_dbl", "pluto_gpu_read_dbl", "pluto_gpu_analysis", "pluto_gpu_write_flt", "pluto_gpu_write_vtk",
]


class PlutoGpuConfig(C.Structure):
    _fields_ = [("dims", C.c_int), ("n", C.c_int * 3), ("recon", C.c_int), ("solver", C.c_int),
                ("rk_order", C.c_int), ("bc", C.c_int * 6), ("arith", C.c_int), ("device", C.c_int),
                ("gamma", C.c_double), ("dx", C.c_double * 3), ("small_dn", C.c_double),
                ("small_pr", C.c_double), ("limiter", C.c_int), ("emf_average", C.c_int),
                ("shock_flattening", C.c_int), ("time_stepping", C.c_int), ("en_correction", C.c_int), ("body_force", C.c_int),
                ("grav", C.c_double * 3), ("char_limiting", C.c_int)]


class PlutoGpuStepInfo(C.Structure):
    _fields_ = [("inv_dt_hyp", C.c_double), ("max_mach", C.c_double),
                ("floor_events", C.c_int), ("nan_events", C.c_int)]


_lib = None


def load_library(path: str | None = None):
    """Load libpluto_gpu.so and declare the prototypes.  Fails loudly."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("PLUTO_GPU_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C pluto_b200/csrc).  pluto_b200 has no CPU fallback.")
    L = C.CDLL(p)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.pluto_gpu_create.argtypes = [C.POINTER(PlutoGpuConfig), C.POINTER(vp)]
    L.pluto_gpu_create.restype = C.c_int
    L.pluto_gpu_destroy.argtypes = [vp]
    L.pluto_gpu_destroy.restype = None
    L.pluto_gpu_last_error.restype = C.c_char_p
    L.pluto_gpu_nghost.argtypes = [vp]
    L.pluto_gpu_nstages.argtypes = [vp]
    L.pluto_gpu_set_body_force.argtypes = [vp, vp, vp, vp]
    L.pluto_gpu_set_grid.argtypes = [vp, vp, vp, vp]
    L.pluto_gpu_set_plm_coeffs.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
    L.pluto_gpu_set_ppm_coeffs.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.pluto_gpu_multi_set_ppm_coeffs.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.pluto_gpu_multi_set_grid.argtypes = [vp, vp, vp, vp]
    L.pluto_gpu_multi_set_plm_coeffs.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
    L.pluto_gpu_multi_set_body_force.argtypes = [vp, vp, vp, vp]
    L.pluto_gpu_multi_set_body_potential.argtypes = [vp, vp, vp, vp, vp]
    L.pluto_gpu_set_body_potential.argtypes = [vp, vp, vp, vp, vp]
    for nm in ("pluto_gpu_upload_interior", "pluto_gpu_download_interior",
               "pluto_gpu_upload_data", "pluto_gpu_download_data"):
        getattr(L, nm).argtypes = [vp, vp, vp, vp, vp]
        getattr(L, nm).restype = C.c_int
    L.pluto_gpu_advance.argtypes = [vp, C.c_double, C.POINTER(PlutoGpuStepInfo)]
    L.pluto_gpu_advance_data.argtypes = [vp, C.c_double, vp, vp, vp, vp, C.POINTER(PlutoGpuStepInfo)]
    L.pluto_gpu_boundary.argtypes = [vp]
    L.pluto_gpu_next_dt.argtypes = [C.c_double] * 4
    L.pluto_gpu_next_dt.restype = C.c_double
    L.pluto_gpu_halo_doubles.argtypes = [vp, C.c_int]
    L.pluto_gpu_halo_doubles.restype = C.c_longlong
    L.pluto_gpu_halo_pack.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.pluto_gpu_halo_unpack.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.pluto_gpu_boundary_dim.argtypes = [vp, C.c_int, C.c_int]
    L.pluto_gpu_step_begin.argtypes = [vp]
    L.pluto_gpu_stage.argtypes = [vp, C.c_int, C.c_double]
    L.pluto_gpu_step_end.argtypes = [vp, C.POINTER(PlutoGpuStepInfo)]
    L.pluto_gpu_stream.argtypes = [vp]
    L.pluto_gpu_stream.restype = vp
    L.pluto_gpu_launch_count.argtypes = [vp]
    L.pluto_gpu_launch_count.restype = C.c_longlong
    L.pluto_gpu_device_bytes.argtypes = [vp]
    L.pluto_gpu_device_bytes.restype = C.c_longlong
    L.pluto_gpu_field.argtypes = [vp, C.c_char_p, C.POINTER(dp), C.POINTER(C.c_longlong * 3),
                                  C.POINTER(C.c_int * 3)]
    L.pluto_gpu_read_field.argtypes = [vp, C.c_char_p, vp]
    L.pluto_gpu_measure_fp64.argtypes = [C.c_int, dp]
    L.pluto_gpu_selftest_arith.argtypes = [C.c_int, C.c_longlong, C.c_ulonglong, C.POINTER(C.c_ulonglong * 3)]
    L.pluto_gpu_halo_nbr_doubles.argtypes = [vp, C.POINTER(C.c_int * 3)]
    L.pluto_gpu_halo_nbr_doubles.restype = C.c_longlong
    L.pluto_gpu_halo_plan.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(vp), C.POINTER(vp)]
    L.pluto_gpu_multi_create.argtypes = [C.POINTER(PlutoGpuConfig), C.POINTER(C.c_int * 3), C.POINTER(C.c_int), C.POINTER(vp)]
    L.pluto_gpu_multi_destroy.argtypes = [vp]
    L.pluto_gpu_multi_destroy.restype = None
    L.pluto_gpu_multi_nghost.argtypes = [vp]
    L.pluto_gpu_multi_nblocks.argtypes = [vp]
    L.pluto_gpu_multi_upload_data.argtypes = [vp, vp, vp, vp, vp]
    L.pluto_gpu_multi_download_data.argtypes = [vp, vp, vp, vp, vp]
    L.pluto_gpu_multi_advance.argtypes = [vp, C.c_double, C.POINTER(PlutoGpuStepInfo)]
    L.pluto_gpu_multi_advance_data.argtypes = [vp, C.c_double, vp, vp, vp, vp, C.POINTER(PlutoGpuStepInfo)]
    L.pluto_gpu_halo_plan_stage.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(vp), C.POINTER(vp)]
    L.pluto_gpu_ipc_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp), C.c_char_p]
    L.pluto_gpu_ipc_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.pluto_gpu_ipc_close.argtypes = [vp]
    L.pluto_gpu_ipc_free.argtypes = [vp]
    L.pluto_gpu_halo_signal.argtypes = [vp, vp, C.c_int, C.POINTER(vp), C.c_ulonglong]
    L.pluto_gpu_halo_wait.argtypes = [vp, vp, C.c_int, vp, C.c_ulonglong]
    L.pluto_gpu_halo_pack_all.argtypes = [vp, C.c_int]
    L.pluto_gpu_halo_unpack_all.argtypes = [vp, C.c_int]
    L.pluto_gpu_halo_pack_all_on.argtypes = [vp, C.c_int, vp]
    L.pluto_gpu_stage_shell.argtypes = [vp, C.c_int, C.c_double]
    L.pluto_gpu_stage_interior.argtypes = [vp, C.c_int]
    L.pluto_gpu_write_dbl.argtypes = [vp, C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_long]
    L.pluto_gpu_read_dbl.argtypes = [vp, C.c_char_p]
    L.pluto_gpu_write_flt.argtypes = [vp, C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_long]
    L.pluto_gpu_write_vtk.argtypes = [vp, C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_long, vp, vp, vp]
    L.pluto_gpu_analysis.argtypes = [vp, dp]
    L.pluto_gpu_set_dt.argtypes = [vp, C.c_double]
    L.pluto_gpu_advance_async.argtypes = [vp, C.c_double, C.c_double]
    L.pluto_gpu_next_dt_async.argtypes = [vp, C.c_double, C.c_double]
    L.pluto_gpu_reduction_slots.argtypes = [vp, C.POINTER(vp)]
    L.pluto_gpu_sync_results.argtypes = [vp, C.c_int, C.POINTER(PlutoGpuStepInfo), dp, C.POINTER(C.c_int), dp]
    L.pluto_gpu_timing.argtypes = [vp, C.c_int]
    L.pluto_gpu_timing_get.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), dp, C.POINTER(C.c_longlong)]
    if path is None:
        _lib = L
    return L


def last_error(L) -> str:
    return (L.pluto_gpu_last_error() or b"").decode()
