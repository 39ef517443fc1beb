"""Block domain decomposition and ghost-zone exchange across GPUs.

Replaces the reference's ArrayLib layer for this path: the Cartesian process
topology (Src/Parallel/al_decompose.c:125-137, Src/initialize.c:83-312), the
per-dimension halo swap AL_Exchange_dim (Src/Parallel/al_exchange_dim.c:58-88)
called at the top of Boundary (Src/boundary.c:98-110) and the MPI_Allreduce(MAX)
of the inverse time step and Mach number (Src/main.c:195-199, 415).

One process per GPU (torch.distributed, NCCL).  Per RK stage and per
dimension, in the order x1 -> x2 -> x3 so that edges and corners are filled:
pack the boundary layers of all cell-centred and staggered fields into two
contiguous device buffers (CUDA kernels behind the C ABI), exchange them with
the two neighbours (NCCL send/recv over NVLink), unpack into the ghost zones,
then apply the physical conditions of that dimension.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from .stepper import GpuStepper, StepInfo
from ._lib import last_error as _last_error

_GRIDS = {
    3: {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2), 16: (2, 2, 4)},
    2: {1: (1, 1, 1), 2: (1, 2, 1), 4: (2, 2, 1), 8: (2, 4, 1), 16: (4, 4, 1)},
}


@dataclass
class BlockLayout:
    """n1 x n2 x n3 rank grid over the global zones; rank = c1 + p1*(c2 + p2*c3)."""
    dims: int
    global_n: tuple
    grid: tuple
    periodic: tuple            # per dimension

    @staticmethod
    def pick_grid(dims, world):
        if world not in _GRIDS[dims]:
            raise ValueError(f"no rank grid for {world} ranks in {dims}-D")
        return _GRIDS[dims][world]

    @classmethod
    def weak(cls, dims, n_per_rank, world, periodic=False):
        grid = cls.pick_grid(dims, world)
        n = list(n_per_rank) + [1] * (3 - len(n_per_rank))
        gn = tuple(n[d] * grid[d] if d < dims else 1 for d in range(3))
        per = (periodic,) * 3 if isinstance(periodic, bool) else tuple(periodic)
        return cls(dims, gn, grid, per)

    @classmethod
    def strong(cls, dims, global_n, world, periodic=False):
        grid = cls.pick_grid(dims, world)
        gn = list(global_n) + [1] * (3 - len(global_n))
        for d in range(dims):
            if gn[d] % grid[d]:
                raise ValueError(f"global zones {gn[d]} not divisible by {grid[d]} ranks in x{d+1}")
        per = (periodic,) * 3 if isinstance(periodic, bool) else tuple(periodic)
        return cls(dims, tuple(gn), grid, per)

    @property
    def world(self):
        return self.grid[0] * self.grid[1] * self.grid[2]

    def coords(self, rank):
        p1, p2, _ = self.grid
        return (rank % p1, (rank // p1) % p2, rank // (p1 * p2))

    def rank_of(self, c):
        p1, p2, _ = self.grid
        return c[0] + p1 * (c[1] + p2 * c[2])

    def local_n(self, rank=0):
        return tuple(self.global_n[d] // self.grid[d] for d in range(3))

    def offset(self, rank):
        c, n = self.coords(rank), self.local_n(rank)
        return tuple(c[d] * n[d] for d in range(3))

    def slice_1d(self, rank, dim, arr, ng):
        """The block's part of a per-direction array of the WHOLE grid (zone widths, reconstruction weights; global_n[dim] + 2 ng
        entries, ghost zones included): its own zones and ng ghost entries on either side."""
        o, n = self.offset(rank)[dim], self.local_n(rank)[dim]
        return np.ascontiguousarray(np.asarray(arr, dtype=np.float64)[o:o + n + 2 * ng])

    def neighbour(self, rank, dim, side):
        """rank of the block abutting `side` (0 low, 1 high) along dim, or None."""
        if self.grid[dim] == 1:
            return None
        c = list(self.coords(rank))
        c[dim] += -1 if side == 0 else 1
        if c[dim] < 0 or c[dim] >= self.grid[dim]:
            if not self.periodic[dim]:
                return None
            c[dim] %= self.grid[dim]
        return self.rank_of(c)

    def neighbours(self, rank):
        """[(offset, rank)] of every face / edge / corner neighbour block (up to 26)."""
        import itertools
        out = []
        c0 = self.coords(rank)
        rng = [(-1, 0, 1) if d < self.dims and self.grid[d] > 1 else (0,) for d in range(3)]
        for o in itertools.product(*rng):
            if o == (0, 0, 0):
                continue
            c, ok = list(c0), True
            for d in range(3):
                c[d] += o[d]
                if c[d] < 0 or c[d] >= self.grid[d]:
                    if not self.periodic[d]:
                        ok = False
                        break
                    c[d] %= self.grid[d]
            if ok:
                out.append((o, self.rank_of(c)))
        return out

    def block_bc(self, rank, physical_bc):
        """bc names of one block: 'shared' where another block abuts (boundary.c:139)."""
        bc = list(physical_bc)
        for d in range(self.dims):
            for side in range(2):
                if self.neighbour(rank, d, side) is not None:
                    bc[2 * d + side] = "shared"
        return tuple(bc)


def exchange_ops_order():
    """Posting order of the four transfers of one dimension.  With two ranks in a
    periodic dimension both neighbours are the SAME peer and NCCL/gloo match
    messages per peer in posting order: my send_lo must meet the peer's recv_hi,
    so sends go (lo, hi) and receives (hi, lo)."""
    return (("send", 0), ("send", 1), ("recv", 1), ("recv", 0))


class HaloExchanger:
    """Per-dimension exchange of packed boundary layers between neighbouring ranks.

    `pack(stage, dim, send_lo, send_hi)` / `unpack(stage, dim, recv_lo, recv_hi)`
    take torch tensors (or None for a side without neighbour); the tensors live
    on whatever device the process group moves (CUDA for NCCL, CPU for gloo)."""

    def __init__(self, layout: BlockLayout, rank: int, halo_doubles, pack, unpack, device, group=None):
        import torch
        self.layout, self.rank, self.pack, self.unpack, self.group = layout, rank, pack, unpack, group
        self.buf = {}
        self.nbr = {}
        for d in range(layout.dims):
            lo, hi = layout.neighbour(rank, d, 0), layout.neighbour(rank, d, 1)
            self.nbr[d] = (lo, hi)
            if lo is None and hi is None:
                continue
            n = int(halo_doubles(d))
            mk = lambda: torch.zeros(n, dtype=torch.float64, device=device)
            self.buf[d] = {"send": [mk() if lo is not None else None, mk() if hi is not None else None],
                           "recv": [mk() if lo is not None else None, mk() if hi is not None else None]}
        self.bytes_per_exchange = sum(sum(t.numel() * 8 for t in b["send"] if t is not None) for b in self.buf.values())

    def exchange_dim(self, stage, dim):
        import torch.distributed as dist
        if dim not in self.buf:
            return
        b, nbr = self.buf[dim], self.nbr[dim]
        self.pack(stage, dim, b["send"][0], b["send"][1])
        ops = []
        for kind, side in exchange_ops_order():
            if nbr[side] is None:
                continue
            if kind == "send":
                ops.append(dist.P2POp(dist.isend, b["send"][side], nbr[side], group=self.group))
            else:
                ops.append(dist.P2POp(dist.irecv, b["recv"][side], nbr[side], group=self.group))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        self.unpack(stage, dim, b["recv"][0], b["recv"][1])


class NeighbourExchanger:
    """All-neighbour exchange: every face, edge and corner neighbour has its own
    send/receive buffer, so the transfers of a stage are independent and go out as
    ONE communication group between one pack and one unpack launch.

    Message matching: with 2 ranks in a periodic dimension several offsets map to
    the same peer and the backends match messages per peer in posting order, so
    sends are posted in ascending offset order and receives in ascending order of
    the SENDER's offset (= minus mine)."""

    def __init__(self, layout: BlockLayout, rank: int, nbr_doubles, plan, pack_all, unpack_all, device, group=None):
        import torch
        self.group = group
        self.pack_all, self.unpack_all = pack_all, unpack_all
        self.nbrs = layout.neighbours(rank)
        mk = lambda o: torch.zeros(int(nbr_doubles(o)), dtype=torch.float64, device=device)
        self.send = [mk(o) for o, _ in self.nbrs]
        self.recv = [mk(o) for o, _ in self.nbrs]
        plan([o for o, _ in self.nbrs], self.send, self.recv)
        self.bytes_per_exchange = sum(t.numel() * 8 for t in self.send)
        self._send_order = sorted(range(len(self.nbrs)), key=lambda q: self.nbrs[q][0])
        self._recv_order = sorted(range(len(self.nbrs)), key=lambda q: tuple(-c for c in self.nbrs[q][0]))

    def post(self):
        """Send the packed buffers / receive the neighbours' ones: one communication group,
        ordered on the CURRENT stream (NCCL) -- nothing is packed or unpacked here."""
        import torch.distributed as dist
        if not self.nbrs:
            return
        ops = [dist.P2POp(dist.isend, self.send[q], self.nbrs[q][1], group=self.group) for q in self._send_order]
        ops += [dist.P2POp(dist.irecv, self.recv[q], self.nbrs[q][1], group=self.group) for q in self._recv_order]
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def exchange(self, stage):
        if not self.nbrs:
            return
        self.pack_all(stage)
        self.post()
        self.unpack_all(stage)


class PeerExchanger:
    """All-neighbour exchange by PEER STORES over NVLink, no communication library on the data path (ranks of one node).

    Every rank owns a receive arena in its HBM: per stage and per neighbour one buffer, plus one 64-bit arrival counter per
    neighbour.  The arenas are exchanged once as CUDA IPC handles; a rank maps its neighbours' arenas and plans its pack
    launch to write THERE (pluto_gpu_halo_plan_stage with the mapped addresses as send buffers): packing is sending.  After the
    pack launch a one-warp kernel stores the exchange number into the neighbours' counters; before the unpack launch a
    one-warp kernel spins until all of this rank's counters have reached it.  The buffer of a stage is reused one step
    later: by then the neighbour has unpacked it (it signalled the exchange in between only after that unpack, in stream
    order), so no acknowledgement travels back.  Steps with a single stage (TIME_STEPPING HANCOCK) would need one: they
    stay with the NCCL exchanger."""

    def __init__(self, layout: BlockLayout, rank: int, block, device: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        if block.nstages < 2:
            raise RuntimeError("peer exchange needs >= 2 stages per step (buffers alternate)")
        self.block, self.L, self.device = block, block.L, device
        self.nbrs = layout.neighbours(rank)
        n = len(self.nbrs)
        self.n = n
        al = lambda b: (b + 255) & ~255
        sizes = [8 * int(block.halo_nbr_doubles(o)) for o, _ in self.nbrs]
        off, self.recv_off = 0, []
        for _ in range(block.nstages):
            row = []
            for q in range(n):
                row.append(off)
                off += al(sizes[q])
            self.recv_off.append(row)
        self.cnt_off = off
        off += al(8 * max(n, 1))
        # every step below is collective: a rank that cannot allocate or map (ranks on different nodes, IPC disabled) must
        # not leave the others waiting, so failures are agreed on before anyone raises
        self.arena = C.c_void_p()
        self.mapped = {}
        handle = C.create_string_buffer(64)
        ok = self.L.pluto_gpu_ipc_alloc(device, max(off, 256), C.byref(self.arena), handle) == 0
        torch.cuda.synchronize()
        mine = dict(ok=ok, handle=handle.raw, recv_off=self.recv_off, cnt_off=self.cnt_off, offsets=[tuple(o) for o, _ in self.nbrs])
        info = [None] * layout.world
        dist.all_gather_object(info, mine, group=group)
        ok = all(i["ok"] for i in info)
        if ok:
            for _, p in self.nbrs:
                if p not in self.mapped:
                    ptr = C.c_void_p()
                    if self.L.pluto_gpu_ipc_open(device, info[p]["handle"], C.byref(ptr)) != 0:
                        ok = False
                        break
                    self.mapped[p] = ptr.value
        flag = torch.tensor([1.0 if ok else 0.0], device=torch.device("cuda", device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if flag.item() != 1.0:
            self.close()
            raise RuntimeError("peer exchange unavailable (CUDA IPC between the ranks): " + _last_error(self.L))
        self.peer_cnt = []
        send = [[] for _ in range(block.nstages)]
        for o, p in self.nbrs:
            qp = info[p]["offsets"].index(tuple(-c for c in o))          # my slot in the neighbour's tables
            self.peer_cnt.append(self.mapped[p] + info[p]["cnt_off"] + 8 * qp)
            for st in range(block.nstages):
                send[st].append(self.mapped[p] + info[p]["recv_off"][st][qp])
        for st in range(block.nstages):
            recv = [self.arena.value + self.recv_off[st][q] for q in range(n)]
            block.halo_plan_stage(st + 1, [o for o, _ in self.nbrs], send[st], recv)
        self.bytes_per_exchange = sum(sizes)
        self.seq_push = self.seq_wait = 0
        dist.barrier(group=group)                                        # every arena is mapped before anyone stores into it

    def push(self, stage, stream_ptr=None):
        """Pack = store into the neighbours' arenas, then signal (both on `stream_ptr` or the block's stream)."""
        if not self.n:
            return
        self.seq_push += 1
        if stream_ptr is None:
            self.block.halo_pack_all(stage)
        else:
            self.block.halo_pack_all_on(stage, stream_ptr)
        self.block.halo_signal(stream_ptr, self.peer_cnt, self.seq_push)

    def wait_unpack(self, stage):
        if not self.n:
            return
        self.seq_wait += 1
        self.block.halo_wait(None, self.n, self.arena.value + self.cnt_off, self.seq_wait)
        self.block.halo_unpack_all(stage)

    def forget_in_flight(self):
        """The state was replaced from outside: an exchange already pushed will never be unpacked."""
        self.seq_wait = self.seq_push

    def close(self):
        for p in self.mapped.values():
            self.L.pluto_gpu_ipc_close(p)
        self.mapped = {}
        if self.arena is not None and self.arena.value:
            self.L.pluto_gpu_ipc_free(self.arena)
        self.arena = None


def agree_step_results(err, out, device, group=None):
    """SUM over the ranks of (failed, floor events per step..., NaN events per step...): every rank raises when any rank
    failed, and every rank returns the global event counts.  `out` = ([dt], [StepInfo], dt_next) of this rank or None."""
    import torch
    import torch.distributed as dist
    from .stepper import PlutoGpuError
    infos = out[1] if out is not None else []
    n = torch.tensor([len(infos)], dtype=torch.int64, device=device)
    dist.all_reduce(n, op=dist.ReduceOp.MAX, group=group)
    nmax = int(n.item())
    v = torch.zeros(1 + 2 * nmax, dtype=torch.float64)
    v[0] = 0.0 if err is None else 1.0
    for q, i in enumerate(infos):
        v[1 + q], v[1 + nmax + q] = i.floor_events, i.nan_events
    v = v.to(device)
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    v = v.tolist()
    if v[0] > 0:
        raise PlutoGpuError(str(err) if err is not None else f"the step failed on {int(v[0])} other rank(s)")
    infos = [StepInfo(i.inv_dt_hyp, i.max_mach, int(v[1 + q]), int(v[1 + nmax + q])) for q, i in enumerate(infos)]
    return out[0], infos, out[2]


class DistStepper:
    """AdvanceStep on one block of a decomposed domain (one process per GPU)."""

    def __init__(self, layout: BlockLayout, rank, dx, recon="plm", solver="hlld", rk_order=2,
                 physical_bc=("periodic",) * 6, gamma=5.0 / 3.0, arith="exact", device=0, exchange="all",
                 overlap=None, **scheme):
        self.layout, self.rank = layout, rank
        self.exchange_mode = exchange       # "all": one 26-neighbour group per stage; "dims": x1->x2->x3 swaps
        # overlap: the exchange of the NEXT stage travels on a second stream while the interior
        # zones of the current stage are completed (all-neighbour exchange only)
        if overlap is None:
            overlap = os.environ.get("PLUTO_GPU_NO_OVERLAP") is None
        self.overlap = bool(overlap) and exchange == "all"
        self._prefetched = None             # stage whose ghost zones are already in flight / received
        self.world = layout.world
        n = layout.local_n(rank)
        self.block = GpuStepper(layout.dims, n, dx, recon=recon, solver=solver, rk_order=rk_order,
                                bc=layout.block_bc(rank, physical_bc), gamma=gamma, arith=arith, device=device,
                                **scheme)          # limiter, emf, flatten, ctu
        self.rk_order = rk_order
        self.nstages = self.block.nstages       # Boundary calls = exchanges per step (1 with TIME_STEPPING HANCOCK)
        self.dims = layout.dims
        self.ex = None
        if self.world > 1:
            import torch
            self._torch = torch
            self._stream = torch.cuda.ExternalStream(self.block.stream, device=torch.device("cuda", device))
            ptr = lambda t: t.data_ptr() if t is not None else None
            self.ex = HaloExchanger(
                layout, rank, self.block.halo_doubles,
                lambda st, d, lo, hi: self.block.halo_pack(st, d, ptr(lo), ptr(hi)),
                lambda st, d, lo, hi: self.block.halo_unpack(st, d, ptr(lo), ptr(hi)),
                device=torch.device("cuda", device))
            self.nex = None
            if exchange == "all":
                b = self.block
                self.nex = NeighbourExchanger(
                    layout, rank, b.halo_nbr_doubles,
                    lambda offs, sb, rb: b.halo_plan(offs, [t.data_ptr() for t in sb], [t.data_ptr() for t in rb]),
                    b.halo_pack_all, b.halo_unpack_all, device=torch.device("cuda", device))
            # ghost zones by peer stores over NVLink (PLUTO_GPU_HALO=peer; all ranks on one node) instead of NCCL send/recv
            self.pex = None
            self.halo = "nccl"
            want = os.environ.get("PLUTO_GPU_HALO", "auto")         # auto: peer stores when the ranks can map each other's memory
            if exchange == "all" and want in ("auto", "peer") and self.block.nstages >= 2:
                try:
                    self.pex = PeerExchanger(layout, rank, self.block, device)
                    self.halo = "peer"
                    # the plan now points at the peers' arenas
                except RuntimeError:
                    if want == "peer":
                        raise
                    self.pex = None                                  # the NCCL plan has to be restored
                    b = self.block
                    b.halo_plan([o for o, _ in self.nex.nbrs], [t.data_ptr() for t in self.nex.send],
                                [t.data_ptr() for t in self.nex.recv])
            self._red = torch.zeros(2, dtype=torch.float64, device=torch.device("cuda", device))
            if self.overlap:
                self._comm = torch.cuda.Stream(device=torch.device("cuda", device))
                self._ev_shell = torch.cuda.Event()
                self._ev_comm = torch.cuda.Event()

    def set_grid(self, *global_dx):
        """Non-uniform grid: the zone widths of the WHOLE grid per direction (ghost zones included); every rank takes its slice."""
        ng = self.block.ng
        self.block.set_grid(*[self.layout.slice_1d(self.rank, d, a, ng) for d, a in enumerate(global_dx[:self.dims])])

    def set_plm_coeffs(self, global_coeffs):
        """UNIFORM_CARTESIAN_GRID NO: the six weight arrays of the WHOLE grid per direction; every rank takes its slice."""
        ng = self.block.ng
        self.block.set_plm_coeffs([[self.layout.slice_1d(self.rank, d, a, ng) for a in six] for d, six in enumerate(global_coeffs)])

    def set_ppm_coeffs(self, global_coeffs):
        """PARABOLIC on a non-uniform grid: the four interface-weight arrays of the WHOLE grid per direction; every rank takes its slice."""
        ng = self.block.ng
        self.block.set_ppm_coeffs([[self.layout.slice_1d(self.rank, d, a, ng) for a in four] for d, four in enumerate(global_coeffs)])

    def _drain(self):
        """The state is about to be replaced from outside: forget the exchange in flight."""
        if self.world > 1 and self.overlap and self._prefetched is not None:
            self._comm.synchronize()
            self._prefetched = None
            if self.pex is not None:
                self.pex.forget_in_flight()

    def set_state(self, dump):
        self._drain()
        self.block.set_state(dump)

    def get_state(self):
        return self.block.get_state()

    def next_dt(self, *a):
        return self.block.next_dt(*a)

    def set_dt(self, dt):
        self.block.set_dt(dt)

    def advance_async(self, cfl, cfl_max_var=1.1):
        """One step with NextTimeStep on the device: nothing here waits for the GPU, the
        all-reduce(MAX) of the CFL reduction runs on the device slots in stream order."""
        if self.world == 1:
            return self.block.advance_async(cfl, cfl_max_var)
        self._enqueue_step(-1.0)
        import torch.distributed as dist
        with self._torch.cuda.stream(self._stream):
            dist.all_reduce(self._red_view(), op=dist.ReduceOp.MAX)
            self.block.next_dt_async(cfl, cfl_max_var)

    def sync_results(self, max_steps=4096):
        """Wait for the enqueued steps.  A failure on ONE rank (Roe solver abort, CUDA error) or its event counts must
        reach every rank -- a rank that raised alone would leave the others waiting in the next exchange -- so the
        failure flag and the floor / NaN counts are summed over the ranks before anything is raised or returned."""
        if self.world == 1:
            return self.block.sync_results(max_steps)
        from .stepper import PlutoGpuError
        err, out = None, None
        try:
            out = self.block.sync_results(max_steps)
        except PlutoGpuError as e:
            err = e
        out = self._agree(err, out)
        return out

    def _agree(self, err, out):
        return agree_step_results(err, out, self._red.device)

    def _red_view(self):
        """torch view of the two device reduction slots (CFL, Mach: non-negative doubles)."""
        if getattr(self, "_redv", None) is None:
            class _Raw:
                pass
            raw = _Raw()
            raw.__cuda_array_interface__ = {"shape": (2,), "typestr": "<f8", "data": (self.block.reduction_slots(), False),
                                            "version": 3, "strides": None}
            self._redv_owner = raw
            self._redv = self._torch.as_tensor(raw, device=self._torch.device("cuda", self.block.cfg.device))
        return self._redv

    def advance(self, dt) -> StepInfo:
        if self.world == 1:
            return self.block.advance(dt)
        torch = self._torch
        import torch.distributed as dist
        b = self.block
        self._enqueue_step(dt)
        from .stepper import PlutoGpuError
        with torch.cuda.stream(self._stream):
            err, info = None, StepInfo(0.0, 0.0, 0, 0)
            try:
                info = b.step_end()
            except PlutoGpuError as e:       # e.g. the Roe solver's abort (roe.c:300-306) on this rank only: every rank
                err = e                      # must still enter the collectives below, then all of them raise
            # MPI_Allreduce(MAX) of invDt_hyp and g_maxMach (main.c:195-199, 415)
            self._red.copy_(torch.tensor([info.inv_dt_hyp, info.max_mach], dtype=torch.float64))
            dist.all_reduce(self._red, op=dist.ReduceOp.MAX)
            r = self._red.tolist()
            _, infos, _ = self._agree(err, ([], [StepInfo(r[0], r[1], info.floor_events, info.nan_events)], 0.0))
        return infos[0]

    def _enqueue_step(self, dt):
        """All stages of one step with their exchanges (dt < 0: the device's own dt)."""
        torch = self._torch
        b = self.block
        with torch.cuda.stream(self._stream):
            b.step_begin()
            for stage in range(1, self.nstages + 1):
                if self.nex is not None and self.overlap:
                    # ghost zones of this stage: already travelling (started while the previous
                    # stage was completing its interior) or exchanged here, in line
                    if self._prefetched == stage:
                        self._stream.wait_event(self._ev_comm)
                    elif self.pex is not None:
                        self.pex.push(stage)
                    else:
                        b.halo_pack_all(stage)
                        self.nex.post()
                    self._prefetched = None
                    if self.pex is not None:
                        self.pex.wait_unpack(stage)
                    else:
                        b.halo_unpack_all(stage)
                    for d in range(self.dims):
                        b.boundary_dim(stage, d)
                    b.stage_shell(stage, dt)
                    self._ev_shell.record(self._stream)
                    nxt = stage + 1 if stage < self.nstages else 1
                    with torch.cuda.stream(self._comm):
                        self._comm.wait_event(self._ev_shell)
                        if self.pex is not None:
                            self.pex.push(nxt, self._comm.cuda_stream)
                        else:
                            b.halo_pack_all_on(nxt, self._comm.cuda_stream)
                            self.nex.post()
                        self._ev_comm.record(self._comm)
                    b.stage_interior(stage)
                    self._prefetched = nxt
                    continue
                if self.pex is not None:
                    self.pex.push(stage)
                    self.pex.wait_unpack(stage)
                    for d in range(self.dims):
                        b.boundary_dim(stage, d)
                elif self.nex is not None:
                    self.nex.exchange(stage)
                    for d in range(self.dims):
                        b.boundary_dim(stage, d)
                else:
                    for d in range(self.dims):
                        self.ex.exchange_dim(stage, d)
                        b.boundary_dim(stage, d)
                b.stage(stage, dt)

    def advance_data(self, dt, Vc, s1, s2, s3=None) -> StepInfo:
        """AdvanceStep on this block's HOST Data arrays: upload, step, download."""
        if self.world == 1:
            return self.block.advance_data(dt, Vc, s1, s2, s3)
        self._drain()
        self.block.upload_data(Vc, s1, s2, s3)
        info = self.advance(dt)
        self.block.download_data(Vc, s1, s2, s3)
        return info


class LocalMultiBlock:
    """Several blocks driven from ONE process (the reference's host is
    single-threaded): the halo buffers of neighbouring blocks are
    handed over directly (same device) -- used by the single-GPU test of the
    decomposition and as the in-process alternative to one process per GPU."""

    def __init__(self, layout: BlockLayout, dx, physical_bc, device=0, exchange="dims", split=False,
                 host_buffers=False, **kw):
        import torch
        self.layout = layout
        # host_buffers: the exchange buffers live in host memory -- only meaningful with the kernel
        # interpreter of tests/emu (lib_path=...), whose "device" pointers are host pointers
        self._sync = (lambda: None) if host_buffers else torch.cuda.synchronize
        self.exchange_mode = exchange
        self.split = split                  # issue every stage as shell + interior (the overlapped form)
        self.blocks = [GpuStepper(layout.dims, layout.local_n(r), dx, bc=layout.block_bc(r, physical_bc),
                                  device=device, **kw) for r in range(layout.world)]
        self.rk_order = self.blocks[0].rk_order
        self.nstages = self.blocks[0].nstages
        dev = torch.device("cpu") if host_buffers else torch.device("cuda", device)
        self.send = {}
        for r, b in enumerate(self.blocks):
            for d in range(layout.dims):
                for side in range(2):
                    if layout.neighbour(r, d, side) is not None:
                        self.send[(r, d, side)] = torch.zeros(b.halo_doubles(d), dtype=torch.float64, device=dev)
        self._torch = torch
        if exchange == "all":
            # send buffer of (rank r, offset o) is the receive buffer of (neighbour, -o)
            self.nsend = {}
            for r, b in enumerate(self.blocks):
                for o, nb in layout.neighbours(r):
                    self.nsend[(r, o)] = torch.zeros(b.halo_nbr_doubles(o), dtype=torch.float64, device=dev)
            for r, b in enumerate(self.blocks):
                nbrs = layout.neighbours(r)
                sp = [self.nsend[(r, o)].data_ptr() for o, _ in nbrs]
                rp = [self.nsend[(nb, tuple(-c for c in o))].data_ptr() for o, nb in nbrs]
                b.halo_plan([o for o, _ in nbrs], sp, rp)

    def set_grid(self, *global_dx):
        for r, b in enumerate(self.blocks):
            b.set_grid(*[self.layout.slice_1d(r, d, a, b.ng) for d, a in enumerate(global_dx[:self.layout.dims])])

    def set_plm_coeffs(self, global_coeffs):
        for r, b in enumerate(self.blocks):
            b.set_plm_coeffs([[self.layout.slice_1d(r, d, a, b.ng) for a in six] for d, six in enumerate(global_coeffs)])

    def set_ppm_coeffs(self, global_coeffs):
        for r, b in enumerate(self.blocks):
            b.set_ppm_coeffs([[self.layout.slice_1d(r, d, a, b.ng) for a in four] for d, four in enumerate(global_coeffs)])

    def set_state(self, global_state):
        lay = self.layout
        for r, b in enumerate(self.blocks):
            o, n = lay.offset(r), lay.local_n(r)
            sl = lambda e1, e2, e3: (slice(o[2], o[2] + n[2] + e3), slice(o[1], o[1] + n[1] + e2), slice(o[0], o[0] + n[0] + e1))
            sub = {}
            for k, v in global_state.items():
                ext = {"Bx1s": (1, 0, 0), "Bx2s": (0, 1, 0), "Bx3s": (0, 0, 1)}.get(k, (0, 0, 0))
                sub[k] = np.ascontiguousarray(v[sl(*ext)])
            b.set_state(sub)

    def get_state(self):
        lay = self.layout
        gn = lay.global_n
        out = None
        for r, b in enumerate(self.blocks):
            st = b.get_state()
            if out is None:
                out = {}
                for k, v in st.items():
                    ext = {"Bx1s": (1, 0, 0), "Bx2s": (0, 1, 0), "Bx3s": (0, 0, 1)}.get(k, (0, 0, 0))
                    out[k] = np.zeros((gn[2] + ext[2], gn[1] + ext[1], gn[0] + ext[0]))
            o, n = lay.offset(r), lay.local_n(r)
            for k, v in st.items():
                out[k][o[2]:o[2] + v.shape[0], o[1]:o[1] + v.shape[1], o[0]:o[0] + v.shape[2]] = v
        return out

    def advance(self, dt) -> StepInfo:
        lay = self.layout
        torch = self._torch
        for b in self.blocks:
            b.step_begin()
        for stage in range(1, self.nstages + 1):
            if self.exchange_mode == "all":
                for b in self.blocks:
                    b.halo_pack_all(stage)
                self._sync()
                for b in self.blocks:
                    b.halo_unpack_all(stage)
                    for d in range(lay.dims):
                        b.boundary_dim(stage, d)
                self._sync()
            for d in range(lay.dims if self.exchange_mode != "all" else 0):
                for r, b in enumerate(self.blocks):
                    lo, hi = self.send.get((r, d, 0)), self.send.get((r, d, 1))
                    if lo is not None or hi is not None:
                        b.halo_pack(stage, d, lo.data_ptr() if lo is not None else None,
                                    hi.data_ptr() if hi is not None else None)
                self._sync()
                for r, b in enumerate(self.blocks):
                    nlo, nhi = lay.neighbour(r, d, 0), lay.neighbour(r, d, 1)
                    rlo = self.send[(nlo, d, 1)].data_ptr() if nlo is not None else None
                    rhi = self.send[(nhi, d, 0)].data_ptr() if nhi is not None else None
                    if rlo is not None or rhi is not None:
                        b.halo_unpack(stage, d, rlo, rhi)
                    b.boundary_dim(stage, d)
                self._sync()
            for b in self.blocks:
                if self.split:
                    b.stage_shell(stage, dt)
                    b.stage_interior(stage)
                else:
                    b.stage(stage, dt)
        infos = [b.step_end() for b in self.blocks]
        return StepInfo(max(i.inv_dt_hyp for i in infos), max(i.max_mach for i in infos),
                        sum(i.floor_events for i in infos), sum(i.nan_events for i in infos))
