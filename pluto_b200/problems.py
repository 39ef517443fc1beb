"""Synthetic initial conditions of the BASELINE.json shapes (numpy, host side).

Each generator returns a state dictionary in the interior (.dbl) layout:
cell-centred ``rho vx1 vx2 [vx3] Bx1 Bx2 [Bx3] prs`` of shape [n3,n2,n1] and
staggered ``Bx1s Bx2s [Bx3s]`` with one more face along their direction.  The
staggered field is the discrete curl of a vector potential sampled on cell
edges, so div B vanishes to round-off, and the cell-centred field is the
face average -- the same construction the reference's start-up uses
(ASSIGN_VECTOR_POTENTIAL, Src/vec_pot_diff.c:79; problem definitions as in
Test_Problems/MHD/{Orszag_Tang,Blast,Rotor}/init.c).  These are benchmark /
property-test inputs; bit-level parity inputs come from the reference's own
dumps (tests/golden).
"""
from __future__ import annotations

import numpy as np

TWO_PI = 6.28318530717959


class Box:
    def __init__(self, dims, n, domain):
        self.dims = dims
        n = list(n) + [1] * (3 - len(n))
        if dims == 2:
            n[2] = 1
        self.n = tuple(int(v) for v in n)
        self.lo = [domain[d][0] for d in range(3)]
        self.hi = [domain[d][1] for d in range(3)]
        self.dx = [(self.hi[d] - self.lo[d]) / self.n[d] for d in range(dims)]

    def centers(self, d, offset=0, count=None):
        """cell centres of direction d for `count` zones starting at zone `offset`"""
        cnt = self.n[d] if count is None else count
        if d >= self.dims:
            return np.zeros(1)
        return self.lo[d] + (np.arange(offset, offset + cnt) + 0.5) * self.dx[d]

    def faces(self, d, offset=0, count=None):
        cnt = self.n[d] if count is None else count
        if d >= self.dims:
            return np.zeros(1)
        return self.lo[d] + np.arange(offset, offset + cnt + 1) * self.dx[d]


def _curl_state(box: Box, prim_fn, A_fn, offset=(0, 0, 0), count=None):
    """Assemble a state from cell-centred primitives and an edge-sampled potential.

    prim_fn(x, y, z) -> dict rho, vx1, vx2, vx3, prs (broadcastable arrays)
    A_fn(x, y, z)    -> (Ax, Ay, Az)
    offset/count select a sub-block (zones) of the global box.
    """
    dims = box.dims
    cnt = list(box.n) if count is None else list(count) + [1] * (3 - len(count))
    if dims == 2:
        cnt[2] = 1
    xc = [box.centers(d, offset[d], cnt[d]) for d in range(3)]
    xf = [box.faces(d, offset[d], cnt[d]) for d in range(3)]
    n1, n2, n3 = cnt
    shp = (n3, n2, n1)

    def grid(ax, ay, az):
        return (ax[None, None, :], ay[None, :, None], az[:, None, None])

    p = prim_fn(*grid(xc[0], xc[1], xc[2]))
    out = {}
    for k in ("rho", "vx1", "vx2", "prs"):
        out[k] = np.ascontiguousarray(np.broadcast_to(np.asarray(p[k], dtype=np.float64), shp))
    if dims == 3:
        out["vx3"] = np.ascontiguousarray(np.broadcast_to(np.asarray(p["vx3"], dtype=np.float64), shp))

    dx = box.dx
    if dims == 2:
        # Az at corners (i+1/2, j+1/2)
        Az = A_fn(*grid(xf[0], xf[1], xc[2]))[2]
        Az = np.broadcast_to(Az, (1, n2 + 1, n1 + 1))
        b1 = (Az[:, 1:, :] - Az[:, :-1, :]) / dx[1]
        b2 = -(Az[:, :, 1:] - Az[:, :, :-1]) / dx[0]
        out["Bx1s"] = np.ascontiguousarray(b1)
        out["Bx2s"] = np.ascontiguousarray(b2)
    else:
        Ax = np.broadcast_to(A_fn(*grid(xc[0], xf[1], xf[2]))[0], (n3 + 1, n2 + 1, n1))
        Ay = np.broadcast_to(A_fn(*grid(xf[0], xc[1], xf[2]))[1], (n3 + 1, n2, n1 + 1))
        Az = np.broadcast_to(A_fn(*grid(xf[0], xf[1], xc[2]))[2], (n3, n2 + 1, n1 + 1))
        out["Bx1s"] = np.ascontiguousarray((Az[:, 1:, :] - Az[:, :-1, :]) / dx[1] - (Ay[1:, :, :] - Ay[:-1, :, :]) / dx[2])
        out["Bx2s"] = np.ascontiguousarray((Ax[1:, :, :] - Ax[:-1, :, :]) / dx[2] - (Az[:, :, 1:] - Az[:, :, :-1]) / dx[0])
        out["Bx3s"] = np.ascontiguousarray((Ay[:, :, 1:] - Ay[:, :, :-1]) / dx[0] - (Ax[:, 1:, :] - Ax[:, :-1, :]) / dx[1])
    out["Bx1"] = 0.5 * (out["Bx1s"][:, :, 1:] + out["Bx1s"][:, :, :-1])
    out["Bx2"] = 0.5 * (out["Bx2s"][:, 1:, :] + out["Bx2s"][:, :-1, :])
    if dims == 3:
        out["Bx3"] = 0.5 * (out["Bx3s"][1:, :, :] + out["Bx3s"][:-1, :, :])
    return out


# ---------------------------------------------------------------------------
def orszag_tang(dims, n, offset=(0, 0, 0), count=None):
    """Orszag-Tang vortex on [0,2pi]^dims, periodic; gamma = 5/3."""
    L = TWO_PI
    box = Box(dims, n, ((0.0, L), (0.0, L), (0.0, L)))
    if dims == 2:
        prim = lambda x, y, z: dict(rho=25.0 / 9.0, prs=5.0 / 3.0, vx1=-np.sin(y) + 0 * x, vx2=np.sin(x) + 0 * y)
        A = lambda x, y, z: (0.0, 0.0, np.cos(y) + 0.5 * np.cos(2.0 * x))
    else:
        c0 = 0.8
        prim = lambda x, y, z: dict(rho=25.0 / 9.0, prs=5.0 / 3.0, vx1=0.0 * x, vx2=-np.sin(z) + 0 * x,
                                    vx3=np.sin(y) + 0 * x)
        A = lambda x, y, z: (c0 * (np.cos(y) + np.cos(2.0 * z)) + 0 * x, c0 * (np.cos(z) - np.cos(x)) + 0 * y,
                             c0 * (-np.cos(y) + np.cos(x)) + 0 * z)
    st = _curl_state(box, prim, A, offset, count)
    return st, dict(dx=box.dx, gamma=5.0 / 3.0, bc=("periodic",) * 6, cfl=0.4 if dims == 2 else 0.3)


def blast(dims, n, p_in=100.0, p_out=1.0, bmag=10.0, theta=45.0, phi=0.0, radius=0.125,
          offset=(0, 0, 0), count=None):
    """MHD blast wave on [-1/2,1/2]^dims, outflow; gamma = 5/3."""
    box = Box(dims, n, ((-0.5, 0.5),) * 3)
    th, ph = np.deg2rad(theta), np.deg2rad(phi)
    B1, B2, B3 = bmag * np.sin(th) * np.cos(ph), bmag * np.sin(th) * np.sin(ph), bmag * np.cos(th)
    if dims == 2:
        B3 = 0.0

    def prim(x, y, z):
        r = np.sqrt(x * x + y * y + (z * z if dims == 3 else 0.0))
        return dict(rho=1.0, vx1=0.0, vx2=0.0, vx3=0.0, prs=np.where(r <= radius, p_in, p_out))

    A = lambda x, y, z: (0.0 * (x + y + z), B3 * x + 0 * (y + z), -B2 * x + B1 * y + 0 * z)
    st = _curl_state(box, prim, A, offset, count)
    return st, dict(dx=box.dx, gamma=5.0 / 3.0, bc=("outflow",) * 6, cfl=0.4 if dims == 2 else 0.3)


def rotor(n, offset=(0, 0, 0), count=None):
    """2-D MHD rotor (Balsara & Spicer) on [-1/2,1/2]^2, outflow; gamma = 1.4."""
    box = Box(2, n, ((-0.5, 0.5),) * 3)
    r0, r1, omega = 0.1, 0.115, 20.0
    Bx = 5.0 / np.sqrt(4.0 * np.pi)

    def prim(x, y, z):
        r = np.sqrt(x * x + y * y)
        f = (r1 - r) / (r1 - r0)
        rs = np.maximum(r, 1e-300)
        rho = np.where(r <= r0, 10.0, np.where(r < r1, 1.0 + 9.0 * f, 1.0))
        vx = np.where(r <= r0, -omega * y, np.where(r < r1, -f * omega * y * r0 / rs, 0.0))
        vy = np.where(r <= r0, omega * x, np.where(r < r1, f * omega * x * r0 / rs, 0.0))
        return dict(rho=rho, vx1=vx, vx2=vy, prs=1.0)

    A = lambda x, y, z: (0.0, 0.0, Bx * y + 0 * x)
    st = _curl_state(box, prim, A, offset, count)
    return st, dict(dx=box.dx, gamma=1.4, bc=("outflow",) * 6, cfl=0.4)


def _splitmix64(seed):
    state = np.uint64(seed)
    mask = (1 << 64) - 1

    def nxt():
        nonlocal state
        s = (int(state) + 0x9E3779B97F4A7C15) & mask
        state = np.uint64(s)
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        return z ^ (z >> 31)
    return nxt


def turbulence(dims, n, seed=20240607, offset=(0, 0, 0), count=None):
    """Decaying-turbulence box on [0,2pi]^dims, periodic:
    integer wave vectors with 1 <= |k|^2 <= 4, amplitudes ~ |k|^-2 with
    splitmix64(seed) deviates and phases, v_rms = B_rms = 1, rho = p = 1."""
    L = TWO_PI
    box = Box(dims, n, ((0.0, L), (0.0, L), (0.0, L)))
    rng = _splitmix64(seed)
    u01 = lambda: (rng() >> 11) * (1.0 / 9007199254740992.0)
    modes = []
    for kz in range(-2, 3):
        for ky in range(-2, 3):
            for kx in range(-2, 3):
                k2 = kx * kx + ky * ky + kz * kz
                if k2 < 1 or k2 > 4:
                    continue
                if kz < 0 or (kz == 0 and ky < 0) or (kz == 0 and ky == 0 and kx < 0):
                    continue
                if dims == 2 and kz != 0:
                    continue
                amp = 1.0 / k2
                av, pv, aa, pa = [0.0] * 3, [0.0] * 3, [0.0] * 3, [0.0] * 3
                for c in range(3):
                    av[c] = amp * (2.0 * u01() - 1.0)
                    pv[c] = 2.0 * np.pi * u01()
                    aa[c] = amp * (2.0 * u01() - 1.0)
                    pa[c] = 2.0 * np.pi * u01()
                if dims == 2:
                    av[2] = 0.0
                    aa[0] = aa[1] = 0.0
                modes.append(((kx, ky, kz), av, pv, aa, pa))
    sv = sum(0.5 * a * a for m in modes for a in m[1])
    sb = 0.0
    for (kx, ky, kz), _, _, (cx, cy, cz), _ in modes:
        sb += 0.5 * ((ky * cz) ** 2 + (kz * cy) ** 2 + (kz * cx) ** 2 + (kx * cz) ** 2 + (kx * cy) ** 2 + (ky * cx) ** 2)
    sv, sb = 1.0 / np.sqrt(sv), 1.0 / np.sqrt(sb)

    def series(x, y, z, comp, amps_idx, ph_idx, scale):
        x, y, z = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), np.asarray(z, dtype=np.float64)
        if x.size * y.size * z.size > (1 << 21) and z.ndim == 3 and z.shape[1:] == (1, 1):
            # large (benchmark) grids: cos(kx x + ky y + kz z + phi) = cos(kz z) cos(a) - sin(kz z) sin(a) with the
            # plane factor a = kx x + ky y + phi, summed per kz first -- 2 outer products per kz instead of one
            # cosine per mode and zone (same field to round-off; the parity fixtures use the direct sum below)
            P, Q = {}, {}
            for m in modes:
                kx, ky, kz = m[0]
                a = m[amps_idx][comp] * scale
                if a == 0.0:
                    continue
                ang = kx * x[0] + ky * y[0] + m[ph_idx][comp]            # [ny or 1, nx or 1]
                P[kz] = P.get(kz, 0.0) + a * np.cos(ang)
                Q[kz] = Q.get(kz, 0.0) - a * np.sin(ang)
            acc = np.zeros(np.broadcast_shapes(x.shape, y.shape, z.shape))
            for kz in P:
                acc += np.cos(kz * z) * P[kz][None]
                acc += np.sin(kz * z) * Q[kz][None]
            return acc
        acc = 0.0
        for m in modes:
            kx, ky, kz = m[0]
            a = m[amps_idx][comp] * scale
            if a == 0.0:
                continue
            acc = acc + a * np.cos(kx * x + ky * y + kz * z + m[ph_idx][comp])
        return acc + 0.0 * (x + y + z)

    def prim(x, y, z):
        return dict(rho=1.0, prs=1.0, vx1=series(x, y, z, 0, 1, 2, sv), vx2=series(x, y, z, 1, 1, 2, sv),
                    vx3=series(x, y, z, 2, 1, 2, sv))

    class _LazyA:                     # _curl_state takes ONE component per call: evaluate only that one
        def __init__(self, x, y, z):
            self.xyz = (x, y, z)

        def __getitem__(self, c):
            return series(*self.xyz, c, 3, 4, sb)

    A = lambda x, y, z: _LazyA(x, y, z)
    st = _curl_state(box, prim, A, offset, count)
    return st, dict(dx=box.dx, gamma=5.0 / 3.0, bc=("periodic",) * 6, cfl=0.4 if dims == 2 else 0.3)


def make(problem, dims, n, **kw):
    if problem == "ot":
        return orszag_tang(dims, n, **kw)
    if problem == "blast":
        return blast(dims, n, **kw)
    if problem == "rotor":
        return rotor(n, **kw)
    if problem == "turb":
        return turbulence(dims, n, **kw)
    raise ValueError(problem)
