#!/usr/bin/env python
"""bench.py -- zone-updates/s of the unsplit Godunov MHD step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--arith exact|fast]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      the reference's CPU path on the host cores

One "step" = one full RK2 AdvanceStep (both stages) of the whole domain; one
zone-update = one interior zone advanced by one step.  At N = 1
the default workload is BASELINE.json configs[1]: MHD blast wave 3-D 256^3,
HLLD + PLM + CT(UCT_CONTACT), RK2, double precision.  For N > 1 every rank
owns a 256^3 block of a larger blast domain (weak scaling) and exchanges
ghost zones with its neighbours every stage.

Prints ONE JSON line (rank 0).  `value` = zones*K / (max over ranks of the
device time of K steps, CUDA events on the library's stream), state resident
in HBM.  `e2e` = the same through the AdvanceStep contract on HOST arrays
(pluto_gpu_advance_data: H2D of Vc,Vs + step + D2H every step, pinned memory).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md 8(d) contract figures per zone-update
ALGO = {
    ("hlld", "plm", 3): dict(bytes=440.0, flops=3300.0),
    ("hlld", "plm", 2): dict(bytes=320.0, flops=1650.0),
    ("roe", "ppm", 2): dict(bytes=320.0, flops=2900.0),
}
FP64_PEAK_TFLOPS_NOMINAL = 37.2       # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz

WORKLOADS = {
    # name: (problem, dims, n per GPU, recon, solver, cfl, first_dt)
    "blast3d_256": ("blast", 3, (256, 256, 256), "plm", "hlld", 0.3, 1e-4),
    "blast3d_128": ("blast", 3, (128, 128, 128), "plm", "hlld", 0.3, 1e-4),
    "turb3d_512": ("turb", 3, (512, 512, 512), "plm", "hlld", 0.3, 1e-3),
    "turb3d_256": ("turb", 3, (256, 256, 256), "plm", "hlld", 0.3, 1e-3),
    "ot3d_256": ("ot", 3, (256, 256, 256), "plm", "hlld", 0.3, 1e-3),
    "ot2d_512": ("ot", 2, (512, 512, 1), "plm", "hlld", 0.4, 1e-3),
    "rotor2d_4096": ("rotor", 2, (4096, 4096, 1), "ppm", "roe", 0.4, 1e-5),
    # development: the other reconstruction / solver pairs at the sizes above (launch-bound A/B, profiles/r2ad_*)
    "blast3d_256_ppm": ("blast", 3, (256, 256, 256), "ppm", "hlld", 0.3, 1e-4),
    "blast3d_256_hllc": ("blast", 3, (256, 256, 256), "plm", "hllc", 0.3, 1e-4),
    "ot3d_256_roe": ("ot", 3, (256, 256, 256), "plm", "roe", 0.3, 1e-3),
    "ot3d_256_ppm_roe": ("ot", 3, (256, 256, 256), "ppm", "roe", 0.3, 1e-3),
    "rotor2d_4096_plm_roe": ("rotor", 2, (4096, 4096, 1), "plm", "roe", 0.4, 1e-5),
    "rotor2d_4096_ppm_hlld": ("rotor", 2, (4096, 4096, 1), "ppm", "hlld", 0.4, 1e-5),
    # strong scaling (BASELINE.json configs[2]): the GLOBAL grid is fixed and split over the ranks
    "ot3d_1024_strong": ("ot", 3, (1024, 1024, 1024), "plm", "hlld", 0.3, 1e-3),
    "ot3d_512_strong": ("ot", 3, (512, 512, 512), "plm", "hlld", 0.3, 1e-3),
}
# per-kernel algorithmic HBM bytes per zone and launch (DESIGN.md "Kernels"):
# sweep: read 8 V + 1 Bn, U (x1: write 5; x2/x3: read 5 + write 5), write 2 face EMFs + 1 sign byte
# fused x1+x2 sweep (FAST): read 8 V + Bx1s + Bx2s, write 5 U + 4 face EMFs + 2 sign bytes (+ C_dt in stage 1) + the 3 (2-D: 1)
# cell-centred EMFs it stores for ct_emf_kernel
SWEEP_BYTES_3D = {"sweep_x1": (9 + 5 + 2) * 8 + 1, "sweep_x2": (9 + 10 + 2) * 8 + 1, "sweep_x3": (9 + 10 + 2) * 8 + 1,
                  "sweep_x1x2": (10 + 5 + 4 + 3) * 8 + 2 + 4}
SWEEP_BYTES_2D = {"sweep_x1": (7 + 4 + 1) * 8 + 1, "sweep_x2": (7 + 8 + 1) * 8 + 1, "sweep_x1x2": (8 + 4 + 2 + 1) * 8 + 2}
# FP64 instructions per launch and zone of the sweep kernels (ncu smsp__sass_thread_inst_executed_op_d*,
# profiles/): what the FP64 pipe, the binding unit of the sweeps, has to issue
SWEEP_FP64_INSTR_3D = {"sweep_x1x2": 2 * 365, "sweep_x3": 365}
# TIME_STEPPING HANCOCK (ctu_kernels.cuh), average of the predictor and the corrector launch of a direction:
# predictor: read 8 V + Bn, write 8 rhs + 2 face EMFs + sign; corrector: read 8 V + Bn + Bn(half) + 16 transverse rhs
# (+ 5 U for x2, x3), write 5 U + 2 face EMFs + sign
CTU_SWEEP_BYTES_3D = {"sweep_x1": ((9 + 8 + 2) * 8 + 1 + (10 + 16 + 5 + 2) * 8 + 1) / 2,
                      "sweep_x2": ((9 + 8 + 2) * 8 + 1 + (10 + 16 + 10 + 2) * 8 + 1) / 2,
                      "sweep_x3": ((9 + 8 + 2) * 8 + 1 + (10 + 16 + 10 + 2) * 8 + 1) / 2}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.p, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        """Start sampling (before the warm-up: nvidia-smi needs ~0.2 s to deliver its first line)."""
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "10"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        """Only samples that arrived inside [t0, t1] (the timed region) are reported."""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            pass
        rows = [r for t, r in self.rows if len(r) >= 7 and (self.t0 is None or self.t0 <= t <= self.t1 + 0.02)]
        if not rows:        # region shorter than the sampling period: fall back to the samples under load around it
            rows = [r for t, r in self.rows if len(r) >= 7 and (self.t0 is None or self.t0 - 0.3 <= t <= self.t1 + 0.3)]
        num = lambda v: v.replace(".", "").isdigit()
        sm = [float(r[0]) for r in rows if num(r[0])]
        mx = [float(r[1]) for r in rows if num(r[1])]
        reasons = set()
        for r in rows:
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        pw = [float(r[2]) for r in rows if num(r[2])]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
#  reference CPU arm / cpu_baseline
# ---------------------------------------------------------------------------
def sample_grid(problem, dims, n):
    """Grid one reference process can hold: 2-D at full size (rotor 1024^2), 3-D at 128^3."""
    if dims == 3:
        return (128, 128, 128)
    if problem == "rotor":
        return (1024, 1024, 1)
    return (min(n[0], 512), min(n[1], 512), 1)


def cpu_reference_run(problem, dims, recon, solver, cfl, first_dt, n_sample, steps, copies, tstep="rk2"):
    """`copies` concurrent serial runs of the compiled reference (oracle/_ref) for `steps` steps on the sample grid,
    wall clock minus a `-maxsteps 0` run of the same set-up (start-up: grid, Init(), vector potential), as BASELINE.md
    4.3 asks.  Returns (zone-updates/s aggregate, wall s of the steps, start-up s, kind)."""
    from oracle.refrun import RefConfig, have_ref, run_reference
    cfg = RefConfig(problem=problem, dims=dims, n=tuple(n_sample), recon=recon, solver=solver,
                    cfl=cfl, first_dt=first_dt, tstep=tstep)
    zones = int(np.prod(n_sample[:dims]))
    if have_ref(cfg):
        def timed(maxsteps):
            th = [threading.Thread(target=run_reference, args=(cfg,), kwargs=dict(maxsteps=maxsteps, no_write=True))
                  for _ in range(copies)]
            t0 = time.perf_counter()
            [t.start() for t in th]
            [t.join() for t in th]
            return time.perf_counter() - t0
        startup = timed(0)
        wall = timed(steps - 1) - startup            # -maxsteps M runs M+1 steps (Src/main.c:133-243)
        return zones * steps * copies / wall, wall, startup, "reference"
    # the compiled reference did not travel: time the CPU restatement instead (one scalar thread)
    from oracle.oracle_lib import Oracle, next_dt
    from pluto_b200 import problems
    st0, meta = problems.make(problem, dims, n_sample)
    o = Oracle(dims, n_sample, meta["dx"], recon=recon, solver=solver, bc=meta["bc"], gamma=meta["gamma"],
               ctu=(tstep == "hancock"))
    o.set_state(st0)
    dt = first_dt
    t0 = time.perf_counter()
    for _ in range(steps):
        inv, _, _ = o.advance(dt)
        dt = next_dt(inv, cfl, 1.1, dt)
    wall = time.perf_counter() - t0
    return zones * steps / wall, wall, 0.0, "port"


def workload_config(name, tstep, world=1):
    """The `config` object both arms print (same workload, same scheme)."""
    problem, dims, n, recon, solver, cfl, first_dt = WORKLOADS[name]
    return {"workload": name, "problem": problem, "dims": dims,
            "zones_per_gpu" if not name.endswith("_strong") else "global_zones": list(n[:dims]),
            "scheme": f"{solver}+{recon}+ct_uct_contact+" + ("ctu_hancock" if tstep == "hancock" else "rk2"),
            "cfl": cfl, "l2": "inputs larger than L2 (state >= 1.5 GB/GPU vs 126 MB L2)" if dims == 3 or n[0] >= 2048
            else "state smaller than L2: a 256 MB buffer is written between timed repetitions"}


def run_reference_arm(args, name):
    """bench.py --impl reference: the compiled reference on the host cores, one serial copy per core (the reference has
    no threads; MPI is not installed), on the sample grid of BASELINE.md 4.2.  A 'step' is one time step of all copies."""
    problem, dims, n, recon, solver, cfl, first_dt = WORKLOADS[name]
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_sample = sample_grid(problem, dims, n)
    copies = os.cpu_count() or 1
    zones = int(np.prod(n_sample[:dims]))
    # bound the run to a few minutes: ~5e5 (3-D) / 1.2e6 (2-D) zone-updates/s per core, slower with every core busy
    est_step_s = zones / (3.5e5 if dims == 3 else 8e5)
    cap = max(2, int(150.0 / est_step_s))
    steps = max(2, min(args.steps if args.steps else 20, cap))
    v, wall, startup, kind = cpu_reference_run(problem, dims, recon, solver, cfl, first_dt, n_sample, steps, copies,
                                               args.time_stepping)
    sample = (f"{copies} concurrent serial copies of the compiled reference (gcc -O3), {problem} {dims}-D "
              f"{'x'.join(str(q) for q in n_sample[:dims])}, {steps} steps each ({wall:.1f} s), start-up run of "
              f"{startup:.1f} s subtracted")
    line = {
        "impl": "reference", "metric": "zone_updates_per_sec", "value": v, "unit": "zone-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 0, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True,
        "scaling": "strong" if name.endswith("_strong") else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(name, args.time_stepping),
        "cpu_baseline": {"value": v, "unit": "zone-updates/s", "cores": copies, "kind": kind, "sample": sample,
                         "per_core": v / copies},
        "e2e": {"value": v, "unit": "zone-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
#  one workload on the GPU(s)
# ---------------------------------------------------------------------------
class Ctx:
    pass


def run_workload(cx, name, steps, warmup, arith, tstep, host_dt=False, want_e2e=True, want_kernels=True, min_region_s=2.0):
    """Time `steps` steps (None: as many as fill min_region_s) of one workload; returns the result dictionary."""
    torch, dist = cx.torch, cx.dist
    from pluto_b200 import problems
    from pluto_b200.parallel import BlockLayout, DistStepper
    problem, dims, n, recon, solver, cfl, first_dt = WORKLOADS[name]
    world, rank, local = cx.world, cx.rank, cx.local
    strong = name.endswith("_strong")
    periodic = problem in ("ot", "turb")
    if strong:       # fixed global grid split over the ranks
        layout = BlockLayout.strong(dims, n, world, periodic=periodic)
        n = layout.local_n(rank)
    else:            # weak scaling: every rank owns one block of n zones of a larger domain
        layout = BlockLayout.weak(dims, n, world, periodic=periodic)
    off = layout.offset(rank)
    st0, meta = problems.make(problem, dims, layout.global_n, offset=off, count=n)
    s = DistStepper(layout, rank, meta["dx"], recon=recon, solver=solver, rk_order=2, physical_bc=meta["bc"],
                    gamma=meta["gamma"], arith=arith, device=local, ctu=(tstep == "hancock"))
    s.set_state(st0)
    del st0
    zones_local = int(np.prod(n[:dims]))
    zones_total = zones_local * world
    stream = torch.cuda.ExternalStream(s.block.stream)
    small = s.block.device_bytes < (300 << 20)          # state that fits the 126 MB L2 (with margin): flush between repetitions
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dt = first_dt
    for _ in range(max(3, warmup)):
        info = s.advance(dt)
        dt = s.next_dt(info.inv_dt_hyp, cfl, 1.1, dt)
    if steps is None:
        # calibrate: the timed region should last >= min_region_s (clock sampling, power state)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, cfl, 1.1, dt)
        barrier()
        per = (time.perf_counter() - t0) / 3
        steps = int(min(4000, max(20, np.ceil(min_region_s / max(per, 1e-6)))))
        if world > 1:
            tt = torch.tensor([steps], dtype=torch.int64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            steps = int(tt.item())

    launches0 = s.block.launch_count
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if small:
        # repetitions of <= 25 steps with an L2 flush (outside the event pairs) between them
        ms, done = 0.0, 0
        s.set_dt(dt)
        while done < steps:
            k = min(25, steps - done)
            flush.fill_(done & 255)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(k):
                s.advance_async(cfl, 1.1)
            e1.record(stream)
            _, infos, dt = s.sync_results()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
            done += k
        info = infos[-1]
    elif host_dt:
        # the reference's loop literally: the host waits for every step's CFL reduction (main.c:133-243)
        e0.record(stream)
        for _ in range(steps):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, cfl, 1.1, dt)
        e1.record(stream)
    else:
        # same dt sequence, NextTimeStep evaluated on the device: the K steps are enqueued back to back
        s.set_dt(dt)
        e0.record(stream)
        for _ in range(steps):
            s.advance_async(cfl, 1.1)
        e1.record(stream)
        _, infos, dt = s.sync_results()
        info = infos[-1]
    barrier()
    t_wall1 = time.time()
    if not small:
        ms = e0.elapsed_time(e1)
    launches = s.block.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    res = {"name": name, "steps": steps, "ms": ms_max, "value": zones_total * steps / (ms_max * 1e-3),
           "ms_per_step": ms_max / steps, "launches": launches, "window": (t_wall0, t_wall1),
           "zones_local": zones_local, "zones_total": zones_total, "layout": layout, "n": n, "dims": dims,
           "device_bytes": s.block.device_bytes, "strong": strong, "nan": info.nan_events,
           "halo": (getattr(s, "halo", None) if world > 1 else None),
           "l2_flush": bool(small)}

    rep, ksteps = None, 0
    if want_kernels:
        # per-kernel device times: a second, shorter pass with CUDA events around every
        # launch (this disables the CUDA-graph replay, so it is kept out of `value`)
        ksteps = max(2, min(steps, 5))
        s.block.timing(True)
        for _ in range(ksteps):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, cfl, 1.1, dt)
        rep = s.block.timing_report()
        s.block.timing(False)
    res["rep"], res["ksteps"] = rep, ksteps

    res["e2e"] = res["e2e_literal"] = None
    if want_e2e:
        bufs = s.block.data_buffers(pinned=True)
        s.block.download_data(*bufs)
        nbytes = sum(b.nbytes for b in bufs if b is not None)
        # (1) the drop-in contract with the state RESIDENT in HBM (integration/advance_step_gpu.c, PLUTO_GPU_RESIDENT=1):
        # host arrays uploaded once, every step driven from the host (dt down, CFL / Mach / event counts back, NextTimeStep
        # on the host), the host arrays refreshed every `sync_every` steps as an output or analysis call would ask
        sync_every, k1 = 10, max(10, min(steps, 50))
        k1 -= k1 % sync_every
        barrier()
        t0 = time.perf_counter()
        s._drain()
        s.block.upload_data(*bufs)
        for q in range(k1):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, cfl, 1.1, dt)
            if (q + 1) % sync_every == 0:
                s.block.download_data(*bufs)
        barrier()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        res["e2e"] = {"value": zones_total * k1 / float(tw.item()), "unit": "zone-updates/s",
                      "h2d_bytes_per_step": nbytes / k1 + 64, "d2h_bytes_per_step": nbytes / sync_every + 24, "steps": k1,
                      "api": "AdvanceStep drop-in with resident state (pluto_gpu_upload_data once, pluto_gpu_advance per step "
                             f"with host NextTimeStep, pluto_gpu_download_data every {sync_every} steps; pinned host Data arrays, "
                             "upload inside the timed region)"}
        # (2) the literal contract: upload + step + download of the whole Data arrays EVERY step (PCIe-bound)
        k2 = max(3, min(steps, 8))
        for _ in range(2):
            s.advance_data(dt, *bufs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k2):
            info = s.advance_data(dt, *bufs)
            dt = s.next_dt(info.inv_dt_hyp, cfl, 1.1, dt)
        barrier()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        res["e2e_literal"] = {"value": zones_total * k2 / float(tw.item()), "unit": "zone-updates/s",
                              "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": k2,
                              "api": "pluto_gpu_advance_data (AdvanceStep on host Data arrays: H2D + step + D2H every step, pinned)"}
        del bufs
    s._drain()
    s.block.close()
    del s
    torch.cuda.empty_cache()
    return res


def rooflines(res, name, tstep, fp64_measured, hbm_peak, peak_src, world):
    """roofline objects of one result: the dominant kernel against the unit that binds it and the whole step against the
    SURVEY.md 8(d) contract figures."""
    problem, dims, n, recon, solver, cfl, first_dt = WORKLOADS[name]
    algo = ALGO.get((solver, recon, dims), dict(bytes=440.0, flops=3300.0))
    step_s = res["ms_per_step"] * 1e-3
    zl = res["zones_local"]
    step_gbs = algo["bytes"] * zl / step_s / 1e9
    step_tf = algo["flops"] * zl / step_s / 1e12
    t_hbm, t_fp = algo["bytes"] / (hbm_peak * 1e9), algo["flops"] / (FP64_PEAK_TFLOPS_NOMINAL * 1e12)
    bound_s = max(t_hbm, t_fp)
    step = {"algorithmic_bytes_per_zone_update": algo["bytes"], "flops_per_zone_update": algo["flops"],
            "binding": "fp64" if t_fp >= t_hbm else "hbm",
            "hbm_gbs": step_gbs, "hbm_frac": step_gbs / hbm_peak,
            "fp64_tflops": step_tf, "fp64_peak_tflops_nominal": FP64_PEAK_TFLOPS_NOMINAL,
            "fp64_frac": step_tf / FP64_PEAK_TFLOPS_NOMINAL,
            "fp64_peak_tflops_measured_dfma_chain": fp64_measured,
            "stencil_roofline_zone_updates_per_sec_per_gpu": 1.0 / bound_s,
            "stencil_roofline_frac": (res["value"] / world) * bound_s}
    rep = res.get("rep")
    if not rep:
        return None, step, None
    ksteps = res["ksteps"]
    top = max((k for k in rep if rep[k][1] > 0), key=lambda k: rep[k][0])
    top_ms, top_cnt = rep[top]
    kern_ms = top_ms / top_cnt
    if tstep == "hancock":
        kern_bytes = CTU_SWEEP_BYTES_3D.get(top, 0) if dims == 3 else 0
    else:
        kern_bytes = (SWEEP_BYTES_3D if dims == 3 else SWEEP_BYTES_2D).get(top, 0)
    hbm_ach = kern_bytes * zl / (kern_ms * 1e-3) / 1e9 if kern_bytes else None
    # contract flops per zone and launch from the SURVEY.md 8(d) hand count: PLM 115 + HLLD 360 + RHS 19 + update 8 + C_dt 4
    # = 506 per direction and stage in 3-D; 370 in 2-D; Roe + PPM 2-D: 2900/4 - 85/2 per direction and stage
    ndir = 2 if top == "sweep_x1x2" else (1 if top.startswith("sweep") else 0)
    per_dir = {("hlld", "plm", 3): 506.0, ("hlld", "plm", 2): 370.0, ("roe", "ppm", 2): 680.0}.get((solver, recon, dims))
    kern_flops = ndir * per_dir if (per_dir and tstep != "hancock") else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tp):      # dram bytes per launch from the committed ncu --set full capture
        traffic = json.load(open(tp)).get(name, {}).get(top)
    step_kernel_ms = sum(v[0] for v in rep.values()) / ksteps
    hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
           "frac": (hbm_ach / hbm_peak) if hbm_ach else None, "algorithmic_bytes_per_zone": kern_bytes,
           "peak_source": peak_src}
    fp64_peak = FP64_PEAK_TFLOPS_NOMINAL
    if kern_flops:
        ach_tf = kern_flops * zl / (kern_ms * 1e-3) / 1e12
        roof = {"bound": "fp64", "kernel": top, "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": ach_tf / fp64_peak, "step_frac": step["stencil_roofline_frac"], "traffic": traffic,
                "peak_source": "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz (SURVEY.md 8d); measured DFMA chain: "
                               + (f"{fp64_measured:.1f} TFLOP/s" if fp64_measured else "n/a"),
                "algorithmic_flops_per_zone": kern_flops,
                "fp64_instr_per_interface_measured": MEASURED_FP64_INSTR.get((solver, recon, dims)),
                "kernel_ms": kern_ms, "kernel_share_of_step": (top_ms / ksteps) / step_kernel_ms,
                "hbm": hbm,
                "note": "the sweeps are bound by the FP64 pipe (SURVEY.md 8d: 89 ps of flops vs 67 ps of bytes per zone-update); "
                        "`frac` = contract flops of this kernel / its launch time / peak, `step_frac` = the whole step against "
                        "the stencil roofline; `hbm` = the same kernel against the measured copy bandwidth (secondary)"}
    else:
        roof = dict(hbm, kernel=top, traffic=traffic, step_frac=step["stencil_roofline_frac"], kernel_ms=kern_ms,
                    kernel_share_of_step=(top_ms / ksteps) / step_kernel_ms)
    kernels = {k: {"ms_per_step": v[0] / ksteps, "launches_per_step": v[1] / ksteps} for k, v in rep.items() if v[1]}
    return roof, step, kernels


def dist_check(cx):
    """N > 1, before anything is timed: every rank advances its 64^3 block of a decomposed periodic turbulence box for 10
    steps with the NCCL exchange (overlapped, device NextTimeStep) and compares it BIT FOR BIT with the same block of a
    single-GPU run of the whole box (the logic of tools/check_dist.py)."""
    torch, dist = cx.torch, cx.dist
    from pluto_b200 import GpuStepper, problems
    from pluto_b200.parallel import BlockLayout, DistStepper
    ok = True
    lay = BlockLayout.weak(3, (64, 64, 64), cx.world, periodic=True)
    gst, meta = problems.make("turb", 3, lay.global_n)
    off, ln = lay.offset(cx.rank), lay.local_n(cx.rank)
    ext = {"Bx1s": (1, 0, 0), "Bx2s": (0, 1, 0), "Bx3s": (0, 0, 1)}
    cut = {k: np.ascontiguousarray(v[off[2]:off[2] + ln[2] + ext.get(k, (0, 0, 0))[2], off[1]:off[1] + ln[1] + ext.get(k, (0, 0, 0))[1],
                                     off[0]:off[0] + ln[0] + ext.get(k, (0, 0, 0))[0]]) for k, v in gst.items()}
    for arith in ("exact", "fast"):
        d = DistStepper(lay, cx.rank, meta["dx"], physical_bc=meta["bc"], gamma=meta["gamma"], device=cx.local, arith=arith)
        one = GpuStepper(3, lay.global_n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"], device=cx.local, arith=arith)
        d.set_state(cut)
        one.set_state(gst)
        dt = 1e-3
        d.set_dt(dt)
        dts = []
        for _ in range(10):
            dts.append(dt)
            a = one.advance(dt)
            d.advance_async(meta["cfl"], 1.1)
            dt = one.next_dt(a.inv_dt_hyp, meta["cfl"], 1.1, dt)
        got, _, dtn = d.sync_results()
        ok = ok and got == dts and dtn == dt
        sa, sb = one.get_state(), d.get_state()
        for k, v in sb.items():
            e = ext.get(k, (0, 0, 0))
            ref = sa[k][off[2]:off[2] + ln[2] + e[2], off[1]:off[1] + ln[1] + e[1], off[0]:off[0] + ln[0] + e[0]]
            ok = ok and bool(np.array_equal(ref, v))
        d._drain()
        d.block.close()
        one.close()
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return "pass" if t.item() == 1.0 else "FAIL"


# FP64-pipe instructions (DFMA + DMUL + DADD + DSETP) per interface solve of the FAST kernels, ncu
# `smsp__sass_thread_inst_executed_op_d*` / source page of profiles/ -- recorded NEXT TO the hand count of SURVEY.md 8(d)
MEASURED_FP64_INSTR = {("hlld", "plm", 3): 339}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: as many as fill a 2 s region)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: blast3d_256 on one GPU (BASELINE configs[1]), turb3d_512 per GPU on several (configs[4])")
    ap.add_argument("--arith", default=os.environ.get("PLUTO_GPU_ARITH", "fast"), choices=["exact", "fast"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_workloads (N = 1), the strong-scaling run and dist_check (N > 1)")
    ap.add_argument("--host-dt", action="store_true", help="host-driven NextTimeStep (a device round trip per step)")
    ap.add_argument("--time-stepping", default="rk2", choices=["rk2", "hancock"],
                    help="hancock: the corner-transport-upwind step (ctu_step.c) instead of RK2")
    ap.add_argument("--dev-lib", default=None, help="development only: time a variant build of libpluto_gpu.so; the line "
                    "carries \"dev_lib\" and is not a bench value")
    args = ap.parse_args()
    if os.environ.get("PLUTO_GPU_LIB"):
        # the test suite may point the loader at the kernel interpreter (tests/emu); a measurement never does
        print("bench.py: PLUTO_GPU_LIB is set -- refusing to time anything but pluto_b200/lib/libpluto_gpu.so", file=sys.stderr)
        sys.exit(2)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    explicit = args.workload is not None
    name = args.workload or ("blast3d_256" if args.gpus == 1 else "turb3d_512")
    if args.impl == "reference":
        run_reference_arm(args, name)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    if args.dev_lib:
        from pluto_b200 import _lib as _pl
        _pl._lib = _pl.load_library(os.path.abspath(args.dev_lib))

    cx = Ctx()
    cx.torch, cx.dist = torch, dist
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.world = world
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if cx.rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run", file=sys.stderr)
        sys.exit(2)
    torch.cuda.set_device(cx.local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", cx.local))
    extras = not args.no_extras and not explicit and not args.dev_lib and args.time_stepping == "rk2"

    check = None
    if world > 1 and not args.no_extras:
        check = dist_check(cx)

    sampler = ClockSampler(cx.local)
    if cx.rank == 0:
        sampler.start()
    res = run_workload(cx, name, args.steps, args.warmup, args.arith, args.time_stepping, host_dt=args.host_dt,
                       want_e2e=not args.no_e2e)
    if cx.rank == 0:
        sampler.window(*res["window"])
    clocks = sampler.stop() if cx.rank == 0 else None
    if res["nan"]:
        print(f"bench.py: rank {cx.rank}: state is not finite", file=sys.stderr)
        sys.exit(3)

    import ctypes
    from pluto_b200 import load_library
    hbm_peak, peak_src = peaks()
    tf = ctypes.c_double(0.0)
    fp64_measured = None
    if cx.rank == 0 and load_library().pluto_gpu_measure_fp64(cx.local, ctypes.byref(tf)) == 0 and tf.value > 0:
        fp64_measured = tf.value

    # ---- the other BASELINE configurations ------------------------------------------------------------------------
    others, strong = None, None
    if extras and world == 1:
        others = {}
        for w in ("turb3d_512", "ot2d_512", "rotor2d_4096"):
            try:
                r = run_workload(cx, w, None, 3, args.arith, "rk2", want_e2e=False, min_region_s=1.0)
                roof, step, kern = rooflines(r, w, "rk2", fp64_measured, hbm_peak, peak_src, 1)
                others[w] = {"value": r["value"], "unit": "zone-updates/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                             "config": workload_config(w, "rk2"), "l2_flush_between_repetitions": r["l2_flush"],
                             "stencil_roofline_frac": step["stencil_roofline_frac"], "binding": step["binding"],
                             "roofline": roof, "kernels": kern, "gpu_launches": r["launches"]}
            except Exception as ex:      # e.g. not enough device memory on a shared box: say so, keep the headline
                others[w] = {"error": str(ex)[:300]}
    if extras and world >= 4:
        try:
            r = run_workload(cx, "ot3d_1024_strong", None, 3, args.arith, "rk2", want_e2e=False, min_region_s=1.0)
            _, step, kern = rooflines(r, "ot3d_1024_strong", "rk2", fp64_measured, hbm_peak, peak_src, world)
            strong = {"workload": "ot3d_1024_strong", "value": r["value"], "unit": "zone-updates/s", "ms_per_step": r["ms_per_step"],
                      "steps": r["steps"], "scaling": "strong", "global_zones": [1024, 1024, 1024],
                      "zones_per_gpu": list(r["n"][:3]), "rank_grid": list(r["layout"].grid),
                      "stencil_roofline_frac": step["stencil_roofline_frac"], "kernels": kern}
        except Exception as ex:
            strong = {"workload": "ot3d_1024_strong", "error": str(ex)[:300]}
    elif extras and world > 1:
        strong = {"workload": "ot3d_1024_strong", "skipped": "1024^3 needs 174 GB per GPU on 2 GPUs (40.4 arrays): run on 4 or 8"}

    if cx.rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    roof, step, kernels = rooflines(res, name, args.time_stepping, fp64_measured, hbm_peak, peak_src, world)
    problem, dims, n, recon, solver, cfl, first_dt = WORKLOADS[name]
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        n_sample = sample_grid(problem, dims, n)
        copies = os.cpu_count() or 1
        ks = 4 if dims == 3 else 30               # ~10-30 s of CPU work per core
        v, wall, startup, kind = cpu_reference_run(problem, dims, recon, solver, cfl, first_dt, n_sample, ks, copies,
                                                   args.time_stepping)
        cpu_baseline = {"value": v, "unit": "zone-updates/s", "cores": copies, "kind": kind, "per_core": v / copies,
                        "sample": f"{copies} concurrent serial copies of the compiled reference (gcc -O3), {problem} {dims}-D "
                                  f"{'x'.join(str(q) for q in n_sample[:dims])}, {ks} steps each, {wall:.1f} s wall, "
                                  f"start-up run of {startup:.1f} s subtracted"}

    cfg = workload_config(name, args.time_stepping, world)
    lay = res["layout"]
    line = {
        "metric": "zone_updates_per_sec", "value": res["value"], "unit": "zone-updates/s", "n_gpus": world,
        "steps": res["steps"], "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if res["strong"] else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "run": {"zones_per_gpu": list(res["n"][:dims]), "global_zones": list(lay.global_n[:dims]), "rank_grid": list(lay.grid),
                "arith": args.arith, "halo": res.get("halo"),
                "next_dt": "host" if args.host_dt else "device kernel, same dt sequence (tests/test_gpu_parity.py)",
                "device_bytes_per_gpu": res["device_bytes"], "timed_region_s": res["ms"] * 1e-3,
                "l2_flush_between_repetitions": res["l2_flush"]},
        "clocks": clocks, "e2e": res["e2e"], "e2e_literal": res["e2e_literal"], "gpu_launches": res["launches"],
        "roofline": roof, "step_roofline": step, "kernels": kernels,
        "cpu_baseline": cpu_baseline,
    }
    if check is not None:
        line["dist_check"] = check
    if others is not None:
        line["other_workloads"] = others
    if strong is not None:
        line["strong"] = strong
    if args.dev_lib:
        line["dev_lib"] = args.dev_lib
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
