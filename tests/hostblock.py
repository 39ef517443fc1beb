"""numpy stand-in for one GPU block's ghost-zone layout (TEST DOUBLE): the same
pack/unpack boxes as pluto_gpu_halo_pack/unpack (pluto_b200/csrc/pluto_gpu.cu,
halo_describe), so that the rank-to-rank exchange logic of
pluto_b200.parallel.HaloExchanger can be exercised on CPU with gloo."""
import numpy as np


class HostBlock:
    def __init__(self, dims, n, ng=2):
        self.dims, self.ng = dims, ng
        self.n = tuple(n)
        self.T = tuple(n[d] + 2 * ng if d < dims else 1 for d in range(3))
        self.beg = tuple(ng if d < dims else 0 for d in range(3))
        self.end = tuple(ng + n[d] - 1 if d < dims else 0 for d in range(3))
        # arrays indexed [k+1][j+1][i+1] like the device layout (index -1 addressable)
        shp = (self.T[2] + 2, self.T[1] + 2, self.T[0] + 2)
        self.nf = (8 if dims == 3 else 6) + dims
        self.f = [np.full(shp, np.nan) for _ in range(self.nf)]
        self.shared_lo = [False, False, False]      # low side of dimension d abuts another block (set by the caller)

    def is_stag(self, q):
        ncell = self.nf - self.dims
        return q - ncell if q >= ncell else None

    def box(self, q, dim, hs, send):
        lo, hi = [0, 0, 0], [self.T[0] - 1, self.T[1] - 1, self.T[2] - 1]
        s = self.is_stag(q)
        if s is not None:
            lo[s] = -1
        normal = s is not None and s == dim
        if send:
            if hs == 0:
                lo[dim], hi[dim] = self.beg[dim], self.beg[dim] + self.ng - 1
            else:
                lo[dim], hi[dim] = self.end[dim] - self.ng + 1 - (1 if normal else 0), self.end[dim]
        else:
            if hs == 0:
                lo[dim], hi[dim] = (-1 if normal else 0), self.beg[dim] - 1
            else:
                lo[dim], hi[dim] = self.end[dim] + 1, self.T[dim] - 1
        return lo, hi

    def view(self, q, lo, hi):
        return self.f[q][lo[2] + 1:hi[2] + 2, lo[1] + 1:hi[1] + 2, lo[0] + 1:hi[0] + 2]

    def halo_doubles(self, dim):
        tot = 0
        for q in range(self.nf):
            lo, hi = self.box(q, dim, 1, True)
            tot += int(np.prod([hi[d] - lo[d] + 1 for d in range(3)]))
        return tot

    def pack(self, stage, dim, send_lo, send_hi):
        for hs, buf in ((0, send_lo), (1, send_hi)):
            if buf is None:
                continue
            off = 0
            b = buf.numpy()
            for q in range(self.nf):
                lo, hi = self.box(q, dim, hs, True)
                v = self.view(q, lo, hi)
                b[off:off + v.size] = v.ravel()
                off += v.size

    def unpack(self, stage, dim, recv_lo, recv_hi):
        for hs, buf in ((0, recv_lo), (1, recv_hi)):
            if buf is None:
                continue
            off = 0
            b = buf.numpy()
            for q in range(self.nf):
                lo, hi = self.box(q, dim, hs, False)
                v = self.view(q, lo, hi)
                v[...] = b[off:off + v.size].reshape(v.shape)
                off += v.size

    def periodic_local(self, dim):
        """local periodic fill of one dimension (reference boundary.c:480-518)"""
        for q in range(self.nf):
            s = self.is_stag(q)
            for hs in (0, 1):
                lo, hi = [0, 0, 0], [self.T[0] - 1, self.T[1] - 1, self.T[2] - 1]
                if s is not None:
                    lo[s] = -1
                if hs == 0:
                    hi[dim] = self.beg[dim] - 1
                else:
                    lo[dim] = self.end[dim] + 1 - (1 if s == dim else 0)
                    lo[dim] = self.end[dim] + 1 if s != dim else self.end[dim]
                src_lo, src_hi = list(lo), list(hi)
                sh = self.n[dim] if hs == 0 else -self.n[dim]
                src_lo[dim] += sh
                src_hi[dim] += sh
                self.view(q, lo, hi)[...] = self.view(q, src_lo, src_hi)


class HostBlockAll(HostBlock):
    """all-neighbour plan (same boxes as nbr_box in pluto_b200/csrc/pluto_gpu.cu)"""

    def nbr_box(self, q, off, send):
        s = self.is_stag(q)
        lo, hi = [0, 0, 0], [0, 0, 0]
        for d in range(3):
            st = 1 if s == d else 0
            if d >= self.dims:
                continue
            if off[d] == 0:        # a shared low face belongs to the low neighbour: it comes with the o[d] = -1 piece
                lo[d], hi[d] = self.beg[d] - (st if not self.shared_lo[d] else 0), self.end[d]
            elif send:
                if off[d] < 0:
                    lo[d], hi[d] = self.beg[d], self.beg[d] + self.ng - 1
                else:
                    lo[d], hi[d] = self.end[d] - self.ng + 1 - st, self.end[d]
            else:
                if off[d] < 0:
                    lo[d], hi[d] = -st, self.beg[d] - 1
                else:
                    lo[d], hi[d] = self.end[d] + 1, self.T[d] - 1
        return lo, hi

    def nbr_doubles(self, off):
        tot = [0, 0]
        for q in range(self.nf):
            for w, send in enumerate((True, False)):
                lo, hi = self.nbr_box(q, off, send)
                tot[w] += int(np.prod([hi[d] - lo[d] + 1 for d in range(3)]))
        return max(tot)

    def plan(self, offsets, send_bufs, recv_bufs):
        self._plan = (list(offsets), send_bufs, recv_bufs)

    def pack_all(self, stage):
        offs, sb, _ = self._plan
        for o, buf in zip(offs, sb):
            pos, b = 0, buf.numpy()
            for q in range(self.nf):
                lo, hi = self.nbr_box(q, o, True)
                v = self.view(q, lo, hi)
                b[pos:pos + v.size] = v.ravel()
                pos += v.size

    def unpack_all(self, stage):
        offs, _, rb = self._plan
        for o, buf in zip(offs, rb):
            pos, b = 0, buf.numpy()
            for q in range(self.nf):
                lo, hi = self.nbr_box(q, o, False)
                v = self.view(q, lo, hi)
                v[...] = b[pos:pos + v.size].reshape(v.shape)
                pos += v.size
