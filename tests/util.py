"""Shared helpers for the parity tests (test infrastructure)."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# BASELINE.json tolerances: per-variable relative L1 error vs the reference
TOL_ONE_STEP = 1e-12
TOL_100_STEPS = 1e-9
TOL_DT = 1e-12


def golden_names(ctu=None):
    """All fixtures, or only those with (ctu=True) / without (ctu=False) corner-transport-upwind time stepping."""
    names = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if ctu is None:
        return names
    return [n for n in names if ("_ctu" in n) == ctu]


class Golden:
    def __init__(self, name):
        d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.problem = str(d["cfg_problem"])
        self.dims = int(d["cfg_dims"])
        n = [int(x) for x in d["cfg_n"]]
        if self.dims == 2:
            n[2] = 1
        self.n = tuple(n)
        self.recon = str(d["cfg_recon"])
        self.solver = str(d["cfg_solver"])
        self.tstep = str(d["cfg_tstep"])
        self.cfl = float(d["cfg_cfl"])
        self.cfl_max_var = float(d["cfg_cfl_max_var"])
        self.first_dt = float(d["cfg_first_dt"])
        self.gamma = float(d["cfg_gamma"])
        self.domain = [tuple(float(x) for x in row) for row in d["cfg_domain"]]
        self.bc = tuple(str(x) for x in d["cfg_bc"])
        self.nsteps = int(d["cfg_nsteps"])
        self.limiter = str(d["cfg_limiter"]) if "cfg_limiter" in d.files else "default"
        self.emf = str(d["cfg_emf"]) if "cfg_emf" in d.files else "uct_contact"
        self.flatten = bool(int(d["cfg_flatten"])) if "cfg_flatten" in d.files else False
        self.en_corr = bool(int(d["cfg_en_corr"])) if "cfg_en_corr" in d.files else False
        self.char_lim = bool(int(d["cfg_char_lim"])) if "cfg_char_lim" in d.files else False
        self.grav = tuple(float(x) for x in d["cfg_grav"]) if "cfg_grav" in d.files else None
        self.grav_mode = int(d["cfg_grav_mode"]) if "cfg_grav_mode" in d.files else 0
        self.potential = bool(int(d["cfg_potential"])) if "cfg_potential" in d.files else False
        self.force = None if self.potential else self.grav      # the `grav=` argument of Oracle / GpuStepper
        # non-uniform grid: the zone widths grid->dx[d] (ghost zones included) and the "Xd-grid" lines that produced them
        self.grid_dx = [d[f"grid_dx{a+1}"] for a in range(self.dims)] if "grid_dx1" in d.files else None
        self.grid = tuple((str(x) or None) for x in d["cfg_grid"]) if "cfg_grid" in d.files else None
        # UNIFORM_CARTESIAN_GRID NO: per direction the six weight arrays of PLM_CoefficientsGet
        self.grid_weights = "cfg_grid_weights" in d.files
        self.plm_coeffs = [list(d[f"plm_coeffs{a+1}"]) for a in range(self.dims)] if "plm_coeffs1" in d.files else None
        # PARABOLIC on a non-uniform grid: per direction the four interface-weight arrays of PPM_CoefficientsGet
        self.ppm_coeffs = [list(d[f"ppm_coeffs{a+1}"]) for a in range(self.dims)] if "ppm_coeffs1" in d.files else None
        self.dt = d["dt"]
        self.states = {}
        for key in d.files:
            if key.startswith("s") and "_" in key and key[1].isdigit():
                s, nm = key.split("_", 1)
                self.states.setdefault(int(s[1:]), {})[nm] = d[key]
        # uniform cell size exactly as the reference computes it
        # (Src/set_grid.c:400  dx = (xR - xL)/npoint)
        self.dx = [(self.domain[a][1] - self.domain[a][0]) / self.n[a] for a in range(self.dims)]
        # what divb_max divides by: the interior zone widths of a non-uniform grid, else the uniform ones
        self.dx_zones = self.dx
        if self.grid_dx is not None:
            self.dx_zones = []
            for a in range(self.dims):
                ngz = (len(self.grid_dx[a]) - self.n[a]) // 2
                self.dx_zones.append(self.grid_dx[a][ngz:ngz + self.n[a]])
            self.dx = [float(np.min(z)) for z in self.dx_zones]      # scale of the smallest zone (tolerances, bscale)
        self.rk_order = 3 if self.tstep == "rk3" else 2
        self.ctu = "chtr" if self.tstep == "chtr" else self.tstep == "hancock"     # the `ctu=` argument of Oracle / GpuStepper


def apply_force_field(stepper, g):
    """Golden fixtures with the static test force (GRAV_MODE 1): hand the per-zone arrays to an Oracle or a GpuStepper.
    Fixtures on a non-uniform grid: hand over the zone widths."""
    if g.grid_dx is not None:
        stepper.set_grid(*g.grid_dx)
    if g.plm_coeffs is not None:
        stepper.set_plm_coeffs(g.plm_coeffs)
    if g.ppm_coeffs is not None:
        stepper.set_ppm_coeffs(g.ppm_coeffs)
    if g.grav_mode == 1:
        stepper.set_body_force(*sign_force_arrays(g.dims, g.n, stepper.ng, g.domain, g.grav, widths=g.grid_dx))
    if g.potential:
        stepper.set_body_potential(*step_potential_arrays(g.dims, g.n, stepper.ng, g.domain, g.grav, widths=g.grid_dx))


def rel_l1(a, b):
    """Per-variable relative L1 error  sum|a-b| / sum|b|  (0 if both vanish)."""
    num = np.abs(a - b).sum()
    den = np.abs(b).sum()
    if den == 0.0:
        return 0.0 if num == 0.0 else np.inf
    return num / den


def max_rel_l1(state, ref):
    return max(rel_l1(state[k], ref[k]) for k in ref)


def divb_max(state, dims, dx):
    """max |sum of face fluxes| / cell volume, from the staggered fields.  dx[d]: the zone width, or the widths of the interior
    zones of a non-uniform grid (Golden.dx_zones)."""
    bx, by = state["Bx1s"], state["Bx2s"]
    w = [np.asarray(dx[d], dtype=float) for d in range(dims)]
    d1 = w[0].reshape(1, 1, -1) if w[0].ndim else w[0]
    d2 = w[1].reshape(1, -1, 1) if w[1].ndim else w[1]
    div = (bx[:, :, 1:] - bx[:, :, :-1]) / d1 + (by[:, 1:, :] - by[:, :-1, :]) / d2
    if dims == 3:
        bz = state["Bx3s"]
        d3 = w[2].reshape(-1, 1, 1) if w[2].ndim else w[2]
        div = div + (bz[1:, :, :] - bz[:-1, :, :]) / d3
    return np.abs(div).max()


def sign_force_arrays(dims, n, ng, domain, grav, widths=None):
    """The static test force of oracle/ref_build/problem/init.c (GRAV_MODE 1): component d = grav[d]*sign(x_d), constant on
    either side of the plane x_d = 0, as arrays [T3][T2][T1] with ghost zones.  Use even n on symmetric domains (no zone
    centre on a plane)."""
    T = [n[d] + 2 * ng if d < dims else 1 for d in range(3)]
    x = []
    for d in range(3):
        if d < dims:
            x.append(_coords(d, n, ng, domain, widths)[0])
        else:
            x.append(np.zeros(1))
    out = []
    for d in range(dims):
        sgn = np.where(x[d] < 0.0, -1.0, 1.0)
        shape = [1, 1, 1]
        shape[2 - d] = T[d]
        out.append(np.ascontiguousarray(np.broadcast_to((grav[d] * sgn).reshape(shape), (T[2], T[1], T[0]))))
    while len(out) < 3:
        out.append(None)
    return out


def _coords(d, n, ng, domain, widths):
    """Zone centres and faces (-1/2 .. T-1/2) of direction d: uniform, or from the zone widths of a non-uniform grid
    (the first interior face lies at the domain's lower end)."""
    T = n[d] + 2 * ng
    if widths is None:
        dx = (domain[d][1] - domain[d][0]) / n[d]
        return domain[d][0] + (np.arange(T) - ng + 0.5) * dx, domain[d][0] + (np.arange(-1, T) - ng + 1.0) * dx
    w = np.asarray(widths[d], dtype=float)
    xf = np.concatenate([[0.0], np.cumsum(w)])
    xf = xf - xf[ng] + domain[d][0]
    return 0.5 * (xf[1:] + xf[:-1]), xf


def step_potential_arrays(dims, n, ng, domain, grav, x0=0.013, widths=None):
    """The test potential of oracle/ref_build/problem/init.c (BODY_FORCE POTENTIAL): steps of height grav[d] across the
    planes x_d = x0, at the zone centres [T3][T2][T1] and at the faces of every direction (staggered Data layouts)."""
    T = [n[d] + 2 * ng if d < dims else 1 for d in range(3)]
    xc, xf = [], []
    for d in range(3):
        if d < dims:
            c_, f_ = _coords(d, n, ng, domain, widths)
            xc.append(c_)
            xf.append(f_)
        else:
            xc.append(np.array([0.5 * (domain[d][0] + domain[d][1])]))
            xf.append(None)

    def phi(x1, x2, x3):
        X3, X2, X1 = np.meshgrid(x3, x2, x1, indexing="ij")
        p = np.zeros(X1.shape)
        p = p + np.where(X1 < x0, grav[0], 0.0)
        p = p + np.where(X2 < x0, grav[1], 0.0)
        p = p + np.where(X3 < x0, grav[2], 0.0)
        return np.ascontiguousarray(p)

    out = [phi(xc[0], xc[1], xc[2])]
    for d in range(dims):
        xs = list(xc)
        xs[d] = xf[d]
        out.append(phi(xs[0], xs[1], xs[2]))
    while len(out) < 4:
        out.append(None)
    return out
