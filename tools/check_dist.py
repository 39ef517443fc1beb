"""Run under torch.distributed.run (one process per GPU): every rank advances its block of a
decomposed domain with NCCL halo exchange and compares it, bit for bit, with the same block cut
out of a single-GPU run of the whole domain."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pluto_b200 import GpuStepper, problems
from pluto_b200.parallel import BlockLayout, DistStepper

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for problem, dims, n, recon, solver in [("ot", 3, (16, 12, 16), "plm", "hlld"), ("blast", 3, (12, 16, 12), "plm", "hlld"),
                                        ("rotor", 2, (24, 20, 1), "ppm", "roe"),
                                        # random-phase field: the blocks' own copies of the periodic faces differ by round-off
                                        ("turb", 3, (12, 12, 16), "plm", "hlld"),
                                        # corner transport upwind: ONE exchange per step, three ghost layers
                                        ("ot", 3, (12, 16, 12), "ctu", "hlld"), ("blast", 2, (24, 20, 1), "ctu", "roe")]:
    periodic = problem in ("ot", "turb")
    ctu = recon == "ctu"
    recon = "plm" if ctu else recon
    lay = BlockLayout.weak(dims, n, world, periodic=periodic)
    gst, meta = problems.make(problem, dims, lay.global_n)
    off, ln = lay.offset(rank), lay.local_n(rank)
    sub, _ = problems.make(problem, dims, lay.global_n, offset=off, count=ln)
    d = DistStepper(lay, rank, meta["dx"], recon=recon, solver=solver, physical_bc=meta["bc"], gamma=meta["gamma"], device=local,
                    ctu=ctu)
    one = GpuStepper(dims, lay.global_n, meta["dx"], recon=recon, solver=solver, bc=meta["bc"], gamma=meta["gamma"], device=local,
                     ctu=ctu)
    # identical inputs: cut the block out of the global arrays (sub-block generation may differ by an ulp)
    cut = {}
    for k, v in gst.items():
        e = {"Bx1s": (1, 0, 0), "Bx2s": (0, 1, 0), "Bx3s": (0, 0, 1)}.get(k, (0, 0, 0))
        cut[k] = np.ascontiguousarray(v[off[2]:off[2] + ln[2] + e[2], off[1]:off[1] + ln[1] + e[1], off[0]:off[0] + ln[0] + e[0]])
    d.set_state(cut); one.set_state(gst)
    dt = {"ot": 5e-3, "blast": 2e-4, "rotor": 1e-3, "turb": 5e-3}[problem]
    for step in range(4):
        a, b = one.advance(dt), d.advance(dt)
        if a.inv_dt_hyp != b.inv_dt_hyp or a.max_mach != b.max_mach:
            ok = False; print(f"rank {rank} {problem}: step {step} scalars differ {a} {b}")
        dt = one.next_dt(a.inv_dt_hyp, meta["cfl"], 1.1, dt)
    # the same with NextTimeStep on the device: steps enqueued back to back, all-reduce on the device slots
    d.set_dt(dt)
    dts = []
    for step in range(4):
        dts.append(dt)
        a = one.advance(dt)
        d.advance_async(meta["cfl"], 1.1)
        dt = one.next_dt(a.inv_dt_hyp, meta["cfl"], 1.1, dt)
    got, _, dtn = d.sync_results()
    if got != dts or dtn != dt:
        ok = False; print(f"rank {rank} {problem}: device dt sequence {got} {dtn} vs host {dts} {dt}")
    sa, sb = one.get_state(), d.get_state()
    for k, v in sb.items():
        e = {"Bx1s": (1, 0, 0), "Bx2s": (0, 1, 0), "Bx3s": (0, 0, 1)}.get(k, (0, 0, 0))
        ref = sa[k][off[2]:off[2] + ln[2] + e[2], off[1]:off[1] + ln[1] + e[1], off[0]:off[0] + ln[0] + e[0]]
        if not np.array_equal(ref, v):
            ok = False; print(f"rank {rank} {problem}: {k} differs, max {np.abs(ref - v).max():.3e}")
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if t.item() == 1.0 else "FAIL", f"world={world}")
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
