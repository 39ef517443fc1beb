"""Host-side decomposition + rank-to-rank ghost exchange on CPU (gloo, 2 and 4
processes): after the x1->x2->x3 exchange every ghost zone, edge and corner of
every field must hold the value of the wrapped global index function."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pluto_b200.parallel import BlockLayout, HaloExchanger, NeighbourExchanger, exchange_ops_order
from tests.hostblock import HostBlock, HostBlockAll


def test_layout_grids_and_neighbours():
    lay = BlockLayout.weak(3, (8, 8, 8), 8, periodic=True)
    assert lay.grid == (2, 2, 2) and lay.global_n == (16, 16, 16)
    assert lay.coords(5) == (1, 0, 1) and lay.rank_of((1, 0, 1)) == 5
    assert lay.offset(5) == (8, 0, 8)
    assert lay.neighbour(0, 0, 0) == 1 and lay.neighbour(0, 0, 1) == 1      # periodic wrap, 2 ranks
    lay = BlockLayout.weak(3, (8, 8, 8), 2, periodic=False)
    assert lay.grid == (1, 1, 2)
    assert lay.neighbour(0, 2, 0) is None and lay.neighbour(0, 2, 1) == 1
    assert lay.block_bc(0, ("outflow",) * 6) == ("outflow",) * 5 + ("shared",)
    assert lay.block_bc(1, ("outflow",) * 6) == ("outflow",) * 4 + ("shared", "outflow")
    lay = BlockLayout.strong(2, (64, 32), 4, periodic=True)
    assert lay.grid == (2, 2, 1) and lay.local_n() == (32, 16, 1)
    with pytest.raises(ValueError):
        BlockLayout.strong(3, (10, 10, 9), 2)
    assert [k for k, _ in exchange_ops_order()] == ["send", "send", "recv", "recv"]


def test_blocks_take_their_slice_of_per_direction_arrays():
    """Zone widths / reconstruction weights of a non-uniform grid are arrays of the whole domain (ghost zones included); a
    block's array is its own zones with ng ghost entries on either side -- the neighbours' zones, or the domain's ghost entries."""
    lay = BlockLayout.strong(3, (16, 12, 8), 8, periodic=False)
    ng = 2
    for d, gn in enumerate((16, 12, 8)):
        arr = np.arange(gn + 2 * ng, dtype=float) + 100 * d
        for rank in range(8):
            o, n = lay.offset(rank)[d], lay.local_n(rank)[d]
            sl = lay.slice_1d(rank, d, arr, ng)
            assert sl.size == n + 2 * ng and sl.flags["C_CONTIGUOUS"]
            assert np.array_equal(sl, arr[o:o + n + 2 * ng])
            assert sl[ng] == arr[ng + o] and sl[-ng - 1] == arr[ng + o + n - 1]      # first / last own zone


def _gfun(q, K, J, I, gn, stag):
    """unique value per (field, wrapped global index); staggered index = face"""
    return q * 1e6 + (K % gn[2]) * 1e4 + (J % gn[1]) * 1e2 + (I % gn[0])


def _worker(rank, world, port, dims, n, periodic, ret, mode="dims", ng=2):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = BlockLayout.weak(dims, n, world, periodic=periodic)
        blk = (HostBlockAll if mode == "all" else HostBlock)(dims, lay.local_n(rank), ng=ng)
        blk.shared_lo = [lay.block_bc(rank, ("periodic",) * 6)[2 * d] == "shared" for d in range(3)]
        off = lay.offset(rank)
        gn = lay.global_n
        ng = blk.ng
        # fill interior zones / interior faces (incl. both boundary faces) only
        idx = [np.arange(-1, blk.T[d] + 1) - blk.beg[d] + off[d] for d in range(3)]   # global index of local -1..T
        for q in range(blk.nf):
            s = blk.is_stag(q)
            K, J, I = np.meshgrid(idx[2], idx[1], idx[0], indexing="ij")
            val = _gfun(q, K, J, I, gn, s)
            lo = [blk.beg[d] for d in range(3)]
            hi = [blk.end[d] for d in range(3)]
            if s is not None:
                lo[s] -= 1
            blk.view(q, lo, hi)[...] = val[lo[2] + 1:hi[2] + 2, lo[1] + 1:hi[1] + 2, lo[0] + 1:hi[0] + 2]
        if mode == "all":
            ex = NeighbourExchanger(lay, rank, blk.nbr_doubles, blk.plan, blk.pack_all, blk.unpack_all, device="cpu")
            ex.exchange(1)
            for d in range(dims):
                if lay.neighbour(rank, d, 0) is None and periodic:
                    blk.periodic_local(d)
        else:
            ex = HaloExchanger(lay, rank, blk.halo_doubles, blk.pack, blk.unpack, device="cpu")
            for d in range(dims):
                ex.exchange_dim(1, d)
                if lay.neighbour(rank, d, 0) is None and periodic:
                    blk.periodic_local(d)
        # every zone (ghosts, edges, corners) must now equal the wrapped global function
        bad = 0
        for q in range(blk.nf):
            s = blk.is_stag(q)
            K, J, I = np.meshgrid(idx[2], idx[1], idx[0], indexing="ij")
            val = _gfun(q, K, J, I, gn, s)
            lo = [0, 0, 0]
            hi = [blk.T[d] - 1 for d in range(3)]
            if s is not None:
                lo[s] = -1
            got = blk.view(q, lo, hi)
            want = val[lo[2] + 1:hi[2] + 2, lo[1] + 1:hi[1] + 2, lo[0] + 1:hi[0] + 2]
            bad += int((got != want).sum())
        # scalar reduction used for dt (MPI_Allreduce MAX, main.c:415)
        t = torch.tensor([float(rank + 1), 10.0 - rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ret[rank] = (bad, t.tolist(), ex.bytes_per_exchange)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["dims", "all"])
@pytest.mark.parametrize("world,dims,n,ng", [(2, 3, (6, 5, 4), 2), (2, 2, (8, 6, 1), 2), (4, 3, (4, 6, 5), 2), (4, 2, (6, 8, 1), 2),
                                             # three ghost layers: PARABOLIC, SHOCK_FLATTENING and the corner-transport-upwind step
                                             (2, 3, (6, 7, 6), 3)])
def test_periodic_exchange_fills_all_ghosts(world, dims, n, ng, mode):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), dims, n, True, ret, mode, ng), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        bad, red, nbytes = ret[r]
        assert bad == 0, f"rank {r}: {bad} ghost values wrong"
        assert red == [float(world), 10.0]
        assert nbytes > 0


def _agree_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pluto_b200.parallel import agree_step_results
        from pluto_b200.stepper import StepInfo, PlutoGpuError
        # (a) nobody failed: the event counts are summed over the ranks, the scalars stay the rank's own (already reduced)
        out = ([1e-3, 2e-3], [StepInfo(5.0, 0.5, rank + 1, 0), StepInfo(6.0, 0.6, 0, 2 * rank)], 3e-3)
        dts, infos, dtn = agree_step_results(None, out, torch.device("cpu"))
        a = (dts, [(i.inv_dt_hyp, i.max_mach, i.floor_events, i.nan_events) for i in infos], dtn)
        # (b) rank 1 alone failed (no results): EVERY rank raises instead of waiting in the next exchange
        raised = False
        try:
            agree_step_results(PlutoGpuError("Roe_Solver: a2 < 0") if rank == 1 else None, None if rank == 1 else out,
                               torch.device("cpu"))
        except PlutoGpuError:
            raised = True
        ret[rank] = (a, raised)
    finally:
        dist.destroy_process_group()


def test_step_failure_and_event_counts_reach_every_rank():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_agree_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in range(2):
        a, raised = ret[r]
        assert a == ([1e-3, 2e-3], [(5.0, 0.5, 3, 0), (6.0, 0.6, 0, 2)], 3e-3)
        assert raised
