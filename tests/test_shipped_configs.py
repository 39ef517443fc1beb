"""Configurations SHIPPED with the reference, unmodified (Test_Problems/MHD/<Problem>/definitions_NN.h, init.c, pluto_NN.ini;
only the output cadence of the ini is changed to one .dbl per step): the reference's own driver with
integration/advance_step_gpu.c + the GPU library must write the same dumps, bit for bit, as the all-CPU reference build.
The binaries are built in the development container by oracle/ref_build/build_shipped.sh from the sources where they lie
(nothing of the reference is stored in the repository) and travel to the GPU box under oracle/_ref/shipped/.

  Orszag_Tang #03  2-D 256^2, LINEAR, RK2, roe, CT_EMF_AVERAGE ARITHMETIC, periodic
  Rotor #01        2-D 400^2, LINEAR, RK2, hlld, ARITHMETIC, MC_LIM, outflow
  Blast #02        3-D 64^3, LINEAR, RK2, roe, VANLEER_LIM, ARITHMETIC, CT_EN_CORRECTION YES, reflective / eqtsymmetric / outflow
  Field_Loop #01   2-D 128 x 64, LINEAR, CHARACTERISTIC_TRACING (corner transport upwind), roe, MC_LIM, UCT_CONTACT, periodic, CFL 0.8
  Field_Loop #02   the same with CT_EMF_AVERAGE UCT0
  Blast #01        2-D 200^2, LINEAR, RK2, roe, VANLEER_LIM, ARITHMETIC, CT_EN_CORRECTION YES, outflow
  Orszag_Tang #05  2-D 256^2, LINEAR, RK3, hlld, MC_LIM, ARITHMETIC, periodic
  Orszag_Tang #09  2-D 512^2, LINEAR, HANCOCK (corner transport upwind), hlld, MC_LIM, ARITHMETIC, CHAR_LIMITING YES, periodic
  Rayleigh_Taylor #05  2-D 256 x 512, LINEAR, HANCOCK (corner transport upwind), roe, MC_LIM, ARITHMETIC, BODY_FORCE VECTOR,
                   periodic / reflective
  Orszag_Tang #07  3-D 64^3, LINEAR, HANCOCK (corner transport upwind), tvdlf, UMIST_LIM, UCT0, periodic
  Rayleigh_Taylor #07  2-D 128 x 256, PARABOLIC, RK3, roe, UCT_CONTACT, BODY_FORCE POTENTIAL, periodic / reflective
"""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from oracle.refrun import read_dbl
from tests.util import ROOT

SHIPPED = os.path.join(ROOT, "oracle", "_ref", "shipped")
CASES = [("orszag_tang_03", 3), ("rotor_01", 2), ("blast_02", 1), ("field_loop_01", 4), ("field_loop_02", 3), ("blast_01", 2),
         ("orszag_tang_05", 2), ("rayleigh_taylor_05", 1), ("orszag_tang_09", 1), ("orszag_tang_07", 1), ("rayleigh_taylor_07", 1)]


def _grid(ini):
    n = []
    for line in open(ini):
        w = line.split()
        if w and w[0] in ("X1-grid", "X2-grid", "X3-grid"):
            n.append(int(w[3]))
    dims = sum(1 for q in n if q > 1)
    return dims, tuple(n)


def _run(binary, ini, nsteps, env=None):
    wd = tempfile.mkdtemp(prefix="shipped_")
    try:
        shutil.copy(ini, os.path.join(wd, "pluto.ini"))
        p = subprocess.run([binary, "-maxsteps", str(nsteps)], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           env=(dict(os.environ, **env) if env else None), timeout=1800)
        assert p.returncode == 0, p.stdout.decode()[-2000:]
        dims, n = _grid(ini)
        dumps = {}
        for line in open(os.path.join(wd, "dbl.out")):
            w = line.split()
            dumps[int(w[3])] = read_dbl(os.path.join(wd, "data.%04d.dbl" % int(w[0])), dims, n)
        return dumps, p.stdout.decode()
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def _compare(tag, nsteps, env):
    ref_bin, gpu_bin, ini = (os.path.join(SHIPPED, tag), os.path.join(SHIPPED, tag + "_gpu"), os.path.join(SHIPPED, tag + ".ini"))
    if not all(os.path.exists(f) for f in (ref_bin, gpu_bin, ini)):
        pytest.skip("oracle/_ref/shipped/* not built (oracle/ref_build/build_shipped.sh)")
    ref, _ = _run(ref_bin, ini, nsteps)
    got, log = _run(gpu_bin, ini, nsteps, env)
    assert "libpluto_gpu" in log
    assert sorted(ref) == sorted(got) and max(ref) >= nsteps
    for s in ref:
        for k, v in ref[s].items():
            assert np.array_equal(got[s][k], v), f"{tag}: {k} differs after {s} steps (max {np.abs(got[s][k]-v).max():.3e})"


@pytest.mark.parametrize("tag,nsteps", CASES)
def test_shipped_configuration_through_the_interpreted_kernels(tag, nsteps, tmp_path):
    """CPU box: the interpreted kernel library (tests/emu) first on the library path of the shim binary."""
    from tests.emu.build_emu import build
    libdir = tmp_path / "lib"
    libdir.mkdir()
    os.symlink(build(), libdir / "libpluto_gpu.so")
    _compare(tag, nsteps, {"PLUTO_GPU_ARITH": "exact", "PLUTO_GPU_NO_GRAPH": "1", "LD_LIBRARY_PATH": str(libdir)})


@pytest.mark.gpu
@pytest.mark.parametrize("tag,nsteps", [(t, 12) for t, _ in CASES])
def test_shipped_configuration_on_the_gpu(tag, nsteps):
    _compare(tag, nsteps, {"PLUTO_GPU_ARITH": "exact"})
