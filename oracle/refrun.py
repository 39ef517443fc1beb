"""Drive the compiled reference binaries under oracle/_ref/ (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke(), tools/make_golden.py and bench.py's
cpu_baseline / --impl reference legs may import this module.  It never
reads /root/reference at run time: the binaries were built from the
reference sources by oracle/ref_build/build_ref.sh and travel with the
repository snapshot.

The on-disk formats read here are the reference's own:
  * ``pluto.ini``   blocks [Grid] [Time] [Solver] [Boundary] [Static Grid
    Output] [Parameters]   (reference Src/runtime_setup.c:22, parse_file.c)
  * ``data.NNNN.dbl`` single_file dumps: for each of rho vx1 vx2 [vx3] Bx1
    Bx2 [Bx3] prs the interior N3xN2xN1 doubles, then Bx1s N3xN2x(N1+1),
    Bx2s N3x(N2+1)xN1, [Bx3s (N3+1)xN2xN1]   (reference Src/bin_io.c:216,
    Src/write_data.c:92-205)
  * ``dt_tap.bin``  (step, t, dt) triples written by the Analysis() tap of
    oracle/ref_build/problem/init.c
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
import time
from dataclasses import dataclass, field

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

PROBLEM_ID = {"ot": 1, "blast": 2, "rotor": 3, "turb": 4}


@dataclass
class RefConfig:
    """One reference run.  Field names follow pluto.ini / definitions.h."""
    problem: str = "ot"                 # ot | blast | rotor | turb
    dims: int = 2
    n: tuple = (64, 64, 1)              # NX1, NX2, NX3
    recon: str = "plm"                  # plm (LINEAR) | ppm (PARABOLIC)
    solver: str = "hlld"                # hlld | hll | roe
    tstep: str = "rk2"                  # rk2 | rk3 | hancock (CTU) | chtr (CTU with TIME_STEPPING CHARACTERISTIC_TRACING)
    limiter: str = "default"            # default | fl mm va os um vl mc  (LIMITER, plm only)
    flatten: bool = False               # SHOCK_FLATTENING MULTID
    emf: str = "uct_contact"            # uct_contact | arith | uct0 | uct_hll  (CT_EMF_AVERAGE)
    char_lim: bool = False              # CHAR_LIMITING YES (plm only; pinned in 2-D, see oracle/mhd_oracle.c states_plm_char)
    grid_weights: bool = False          # UNIFORM_CARTESIAN_GRID NO: reconstruction weights from the grid (plm_coeffs.c)
    en_corr: bool = False               # CT_EN_CORRECTION YES
    grav: tuple = None                  # BODY_FORCE VECTOR with the uniform acceleration (g1, g2, g3)
    grav_mode: int = 0                  # 1: static position-dependent force, component d = grav[d]*sign(x_d)
    potential: bool = False             # BODY_FORCE POTENTIAL: step potential of height grav[d] across x_d = 0.013
    vector_too: bool = False            # with potential: BODY_FORCE (VECTOR+POTENTIAL), the uniform acceleration grav as well
    cfl: float = 0.4
    cfl_max_var: float = 1.1
    first_dt: float = 1.0e-3
    tstop: float = 1.0e10
    gamma: float = 5.0 / 3.0
    domain: tuple = None                # ((x1b,x1e),(x2b,x2e),(x3b,x3e)); default per problem
    bc: tuple = None                    # 6 strings; default per problem
    grid: tuple = None                  # non-uniform grid: per direction the text of the "Xd-grid" line after the keyword, e.g.
                                        # "2  -0.5  12  u  0.0  20  s  0.5" (patches: set_grid.c:330-560), or None: one uniform patch
    blast: dict = field(default_factory=lambda: dict(
        P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=0.0, RADIUS=0.125))
    seed: int = 20240607
    prefix: str = "pluto_"              # "pluto_gpu_" = reference driver + integration/advance_step_gpu.c

    def variant(self) -> str:
        v = f"{self.dims}d_{self.recon}"
        if self.tstep == "rk3":
            v += "_rk3"
        if self.tstep == "hancock":         # CTU with the MUSCL-Hancock predictor (ctu_step.c, hancock.c)
            v += "_hancock"
        if self.tstep == "chtr":            # CTU with the characteristic-tracing predictor (ctu_step.c, char_tracing.c)
            v += "_chtr"
        if self.limiter != "default":
            v += "_l" + self.limiter
        if self.emf != "uct_contact":
            v += "_e" + self.emf
        if self.flatten:
            v += "_sfl"
        if self.en_corr:
            v += "_en"
        if self.char_lim:
            v += "_cl"
        if self.grid_weights:
            v += "_nuw"
        if self.grav is not None:
            v += ("_bfp" if self.vector_too else "_bp") if self.potential else "_bf"
        return v

    def binary(self) -> str:
        return os.path.join(REF_DIR, self.prefix + self.variant())

    def resolved_domain(self):
        if self.domain is not None:
            return self.domain
        if self.problem in ("ot", "turb"):
            L = 6.28318530717959
            return ((0.0, L), (0.0, L), (0.0, L))
        return ((-0.5, 0.5), (-0.5, 0.5), (-0.5, 0.5))

    def resolved_bc(self):
        if self.bc is not None:
            return self.bc
        if self.problem in ("ot", "turb"):
            return ("periodic",) * 6
        return ("outflow",) * 6

    def resolved_gamma(self):
        if self.problem == "rotor":
            return 1.4
        return self.gamma


def have_ref(cfg: RefConfig) -> bool:
    return os.path.isfile(cfg.binary()) and os.access(cfg.binary(), os.X_OK)


def write_ini(cfg: RefConfig, path: str, dbl_dn: int = -1, analysis_dn: int = 1, flt_dn: int = -1, vtk_dn: int = -1):
    dom = cfg.resolved_domain()
    bc = cfg.resolved_bc()
    n = list(cfg.n)
    if cfg.dims == 2:
        n[2] = 1
    lines = ["[Grid]", ""]
    for d in range(3):
        lo, hi = dom[d]
        if cfg.dims == 2 and d == 2:
            lo, hi = 0.0, 1.0
        if cfg.grid is not None and d < len(cfg.grid) and cfg.grid[d] and not (cfg.dims == 2 and d == 2):
            lines.append(f"X{d+1}-grid    {cfg.grid[d]}")
        else:
            lines.append(f"X{d+1}-grid    1    {lo!r}    {n[d]}    u    {hi!r}")
    lines += ["", "[Chombo Refinement]", "", "Levels           4",
              "Ref_ratio        2 2 2 2 2", "Regrid_interval  2 2 2 2",
              "Refine_thresh    0.3", "Tag_buffer_size  3", "Block_factor     4",
              "Max_grid_size    32", "Fill_ratio       0.75", "",
              "[Time]", "",
              f"CFL              {cfg.cfl!r}",
              f"CFL_max_var      {cfg.cfl_max_var!r}",
              f"tstop            {cfg.tstop!r}",
              f"first_dt         {cfg.first_dt!r}", "",
              "[Solver]", "", f"Solver         {cfg.solver}", "",
              "[Boundary]", ""]
    names = ["X1-beg", "X1-end", "X2-beg", "X2-end", "X3-beg", "X3-end"]
    for nm, b in zip(names, bc):
        lines.append(f"{nm}        {b}")
    lines += ["", "[Static Grid Output]", "", "uservar    0",
              f"dbl       -1.0  {dbl_dn}   single_file",
              f"flt       -1.0  {flt_dn}   single_file",
              f"vtk       -1.0  {vtk_dn}   single_file",
              "tab       -1.0  -1   ", "ppm       -1.0  -1   ",
              "png       -1.0  -1   ", "log        100000 ",
              f"analysis  -1.0  {analysis_dn} ", "",
              "[Chombo HDF5 output]", "", "Checkpoint_interval  -1.0  0",
              "Plot_interval         1.0  0", "",
              "[Parameters]", ""]
    b = cfg.blast
    params = [("PROBLEM", PROBLEM_ID[cfg.problem]), ("GAMMA_EOS", cfg.resolved_gamma()),
              ("P_IN", b["P_IN"]), ("P_OUT", b["P_OUT"]), ("BMAG", b["BMAG"]),
              ("THETA", b["THETA"]), ("PHI", b["PHI"]), ("RADIUS", b["RADIUS"]),
              ("SEED", cfg.seed)]
    gr = cfg.grav if cfg.grav is not None else (0.0, 0.0, 0.0)
    params += [("GRAV1", gr[0]), ("GRAV2", gr[1]), ("GRAV3", gr[2]), ("GRAV_MODE", cfg.grav_mode)]
    for k, v in params:
        lines.append(f"{k:<26s}  {float(v)!r}  ")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def field_names(dims: int):
    if dims == 2:
        return ["rho", "vx1", "vx2", "Bx1", "Bx2", "prs"], ["Bx1s", "Bx2s"]
    return ["rho", "vx1", "vx2", "vx3", "Bx1", "Bx2", "Bx3", "prs"], ["Bx1s", "Bx2s", "Bx3s"]


def read_dbl(path: str, dims: int, n):
    """Return dict name -> array [k][j][i] (interior; staggered +1 face)."""
    n1, n2, n3 = n[0], n[1], (n[2] if dims == 3 else 1)
    raw = np.fromfile(path, dtype="<f8")
    cc, st = field_names(dims)
    out = {}
    off = 0
    for nm in cc:
        sz = n1 * n2 * n3
        out[nm] = raw[off:off + sz].reshape(n3, n2, n1).copy()
        off += sz
    shapes = {"Bx1s": (n3, n2, n1 + 1), "Bx2s": (n3, n2 + 1, n1), "Bx3s": (n3 + 1, n2, n1)}
    for nm in st:
        shp = shapes[nm]
        sz = shp[0] * shp[1] * shp[2]
        out[nm] = raw[off:off + sz].reshape(shp).copy()
        off += sz
    if off != raw.size:
        raise ValueError(f"{path}: size mismatch ({raw.size} doubles, consumed {off})")
    return out


@dataclass
class RefResult:
    dumps: dict            # step number -> dict of arrays
    dt_tap: np.ndarray     # rows (step, t, dt) for steps >= 1
    wall_s: float
    steps_run: int
    workdir: str
    stdout: str = ""       # the driver's log (serial build: print() goes to stdout)
    dx: list = None        # zone widths grid->dx[d] (ghost zones included) from grid_tap.bin, one array per direction
    plm_coeffs: list = None  # UNIFORM_CARTESIAN_GRID NO builds: per direction [cp, cm, wp, wm, dp, dm] (PLM_CoefficientsGet)
    ppm_coeffs: list = None  # PARABOLIC builds: per direction the interface weights [w(-1), w(0), w(1), w(2)] of every zone (PPM_CoefficientsGet)


def run_reference(cfg: RefConfig, maxsteps: int, dump_every: int = -1,
                  workdir: str | None = None, keep: bool = False,
                  no_write: bool = False, timeout: float = 3600.0, env: dict | None = None,
                  analysis_every: int = 1, flt_every: int = -1, vtk_every: int = -1) -> RefResult:
    """Run ``pluto -maxsteps M``.

    Reference main loop semantics (Src/main.c:133-243): ``-maxsteps M`` with
    M >= 1 executes M+1 steps (steps 0..M), M == 0 executes none.  Dumps with
    ``dbl -1.0 dn`` are written by CheckForOutput at the top of iteration s
    when s % dn == 0 (1 <= s <= M-1... and not on the last iteration), i.e.
    they hold the state after s steps; data.0000.dbl is the initial state and
    the last file is the final state after M+1 steps.
    """
    if not have_ref(cfg):
        raise FileNotFoundError(cfg.binary())
    own = workdir is None
    if own:
        workdir = tempfile.mkdtemp(prefix="plutoref_")
    os.makedirs(workdir, exist_ok=True)
    write_ini(cfg, os.path.join(workdir, "pluto.ini"), dbl_dn=dump_every,
              analysis_dn=analysis_every, flt_dn=flt_every, vtk_dn=vtk_every)
    cmd = [cfg.binary(), "-maxsteps", str(maxsteps)]
    if no_write:
        cmd.append("-no-write")
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=timeout,
                       env=(dict(os.environ, **env) if env else None))
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("reference run failed:\n" + p.stdout.decode()[-2000:])
    dumps = {}
    steps_run = maxsteps + 1 if maxsteps >= 1 else 0
    if not no_write:
        # dbl.out (reference Src/write_data.c:365-395): one line per dump,
        # "<nfile> <t> <dt> <nstep> single_file little <names...>"
        with open(os.path.join(workdir, "dbl.out")) as f:
            for line in f:
                w = line.split()
                if len(w) < 4:
                    continue
                nfile, nstep = int(w[0]), int(w[3])
                dumps[nstep] = read_dbl(
                    os.path.join(workdir, "data.%04d.dbl" % nfile), cfg.dims, cfg.n)
    tap_path = os.path.join(workdir, "dt_tap.bin")
    tap = (np.fromfile(tap_path, dtype="<f8").reshape(-1, 3)
           if os.path.exists(tap_path) else np.zeros((0, 3)))
    dx = plmc = ppmc = None
    gpath = os.path.join(workdir, "grid_tap.bin")
    if os.path.exists(gpath):
        raw = np.fromfile(gpath, dtype="<f8")
        dx, plmc, ppmc, off = [], [], [], 0
        while off < raw.size:
            m = int(raw[off])
            if m == -4:                # four interface-weight arrays of the next direction (PARABOLIC: wp[i][-1 .. 2])
                t = len(dx[len(ppmc)])
                ppmc.append([raw[off + 1 + q * t:off + 1 + (q + 1) * t].copy() for q in range(4)])
                off += 1 + 4 * t
                continue
            if m == -6:                # six reconstruction-weight arrays of the next direction (UNIFORM_CARTESIAN_GRID NO)
                t = len(dx[len(plmc)])
                plmc.append([raw[off + 1 + q * t:off + 1 + (q + 1) * t].copy() for q in range(6)])
                off += 1 + 6 * t
                continue
            dx.append(raw[off + 1:off + 1 + m].copy())
            off += 1 + m
    res = RefResult(dumps=dumps, dt_tap=tap, wall_s=wall, steps_run=steps_run,
                    workdir=workdir, stdout=p.stdout.decode(errors="replace"), dx=dx, plm_coeffs=(plmc or None),
                    ppm_coeffs=(ppmc or None))
    if own and not keep:
        shutil.rmtree(workdir, ignore_errors=True)
    return res
