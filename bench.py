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
    # strong scaling (BASELINE.json configs[2]): the GLOBAL grid is fixed and split over the ranks
    "ot3d_1024_strong": ("ot", 3, (1024, 1024, 1024), "plm", "hlld", 0.3, 1e-3),
    "ot3d_512_strong": ("ot", 3, (512, 512, 512), "plm", "hlld", 0.3, 1e-3),
}
# per-kernel algorithmic HBM bytes per zone and launch (DESIGN.md "Kernels"):
# sweep: read 8 V + 1 Bn, U (x1: write 5; x2/x3: read 5 + write 5), write 2 face EMFs + 1 sign byte
# fused x1+x2 sweep (FAST): read 8 V + Bx1s + Bx2s, write 5 U + 4 face EMFs + 2 sign bytes (+ C_dt in stage 1)
SWEEP_BYTES_3D = {"sweep_x1": (9 + 5 + 2) * 8 + 1, "sweep_x2": (9 + 10 + 2) * 8 + 1, "sweep_x3": (9 + 10 + 2) * 8 + 1,
                  "sweep_x1x2": (10 + 5 + 4) * 8 + 2 + 4}
SWEEP_BYTES_2D = {"sweep_x1": (7 + 4 + 1) * 8 + 1, "sweep_x2": (7 + 8 + 1) * 8 + 1, "sweep_x1x2": (8 + 4 + 2) * 8 + 2}
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
def cpu_reference_run(problem, dims, recon, solver, cfl, first_dt, n_sample, steps, copies, tstep="rk2"):
    """Time `copies` concurrent runs of the compiled reference (oracle/_ref) on a
    bounded sample grid.  Returns (zone-updates/s aggregate, wall seconds, kind)."""
    from oracle.refrun import RefConfig, have_ref, run_reference
    cfg = RefConfig(problem=problem, dims=dims, n=tuple(n_sample), recon=recon, solver=solver,
                    cfl=cfl, first_dt=first_dt, tstep=tstep)
    zones = int(np.prod(n_sample[:dims]))
    if have_ref(cfg):
        res = [None] * copies

        def work(q):
            res[q] = run_reference(cfg, maxsteps=steps - 1, no_write=True)

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(q,)) for q in range(copies)]
        [t.start() for t in th]
        [t.join() for t in th]
        wall = time.perf_counter() - t0
        return zones * steps * copies / wall, wall, "reference"
    # the compiled reference did not travel: time the CPU restatement instead
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
    return zones * steps / wall, wall, "port"


def sample_plan(dims, steps_hint=None):
    """Bounded CPU sample: ~10-30 s per copy at ~5e5 (3-D) / 1.2e6 (2-D) zone-updates/s/core."""
    if dims == 3:
        return (64, 64, 64), 24          # 6.3e6 zone-updates  ~ 11-15 s per copy
    return (512, 512, 1), 40              # 1.0e7 zone-updates  ~ 9-12 s per copy


def run_reference_arm(args, wl):
    problem, dims, n, recon, solver, cfl, first_dt = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample, steps_per = sample_plan(dims)
    copies = os.cpu_count() or 1
    # each "step" of this arm = one bounded sample; keep the whole run within minutes
    vals = []
    total = max(1, min(args.steps, 3))
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(problem, dims, recon, solver, cfl, first_dt, n_sample, 2, copies, args.time_stepping)
    t0 = time.perf_counter()
    for _ in range(total):
        v, wall, kind = cpu_reference_run(problem, dims, recon, solver, cfl, first_dt, n_sample, steps_per, copies,
                                          args.time_stepping)
        vals.append(v)
    value = float(np.mean(vals))
    sample = (f"{copies} concurrent serial copies of the compiled reference, {problem} {dims}-D "
              f"{'x'.join(str(v) for v in n_sample[:dims])}, {steps_per} steps each, x{total}")
    line = {
        "impl": "reference", "metric": "zone_updates_per_sec", "value": value, "unit": "zone-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (time.perf_counter() - t0) / total, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload,
                   "scheme": f"{solver}+{recon}+ct_uct_contact+" + ("ctu_hancock" if args.time_stepping == "hancock" else "rk2"),
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "zone-updates/s", "cores": copies, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "zone-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="blast3d_256", choices=sorted(WORKLOADS))
    ap.add_argument("--arith", default=os.environ.get("PLUTO_GPU_ARITH", "fast"), choices=["exact", "fast"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from pluto_b200 import problems
    from pluto_b200.parallel import BlockLayout, DistStepper
    if args.dev_lib:
        from pluto_b200 import _lib as _pl
        _pl._lib = _pl.load_library(os.path.abspath(args.dev_lib))

    problem, dims, n, recon, solver, cfl, first_dt = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run", file=sys.stderr)
        sys.exit(2)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    strong = args.workload.endswith("_strong")
    if strong:       # fixed global grid split over the ranks
        layout = BlockLayout.strong(dims, n, world, periodic=(problem in ("ot", "turb")))
        n = layout.local_n(rank)
    else:            # weak scaling: every rank owns one block of n zones of a larger domain
        layout = BlockLayout.weak(dims, n, world, periodic=(problem in ("ot", "turb")))
    off = layout.offset(rank)
    st0, meta = problems.make(problem, dims, layout.global_n, offset=off, count=n)
    s = DistStepper(layout, rank, meta["dx"], recon=recon, solver=solver, rk_order=2, physical_bc=meta["bc"],
                    gamma=meta["gamma"], arith=args.arith, device=local, ctu=(args.time_stepping == "hancock"))
    s.set_state(st0)
    del st0
    zones_local = int(np.prod(n[:dims]))
    zones_total = zones_local * world
    stream = torch.cuda.ExternalStream(s.block.stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dt = first_dt
    for _ in range(args.warmup):
        info = s.advance(dt)
        dt = s.next_dt(info.inv_dt_hyp, cfl, 1.1, dt)

    launches0 = s.block.launch_count
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.host_dt:
        # the reference's loop literally: the host waits for every step's CFL reduction (main.c:133-243)
        e0.record(stream)
        for _ in range(args.steps):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, cfl, 1.1, dt)
        e1.record(stream)
    else:
        # same dt sequence, NextTimeStep evaluated on the device: the K steps are enqueued back to back
        s.set_dt(dt)
        e0.record(stream)
        for _ in range(args.steps):
            s.advance_async(cfl, 1.1)
        e1.record(stream)
        _, infos, dt = s.sync_results()
        info = infos[-1]
    barrier()
    sampler.window(t_wall0, time.time())
    ms = e0.elapsed_time(e1)
    launches = s.block.launch_count - launches0
    # per-kernel device times: a second, shorter pass with CUDA events around every
    # launch (this disables the CUDA-graph replay, so it is kept out of `value`)
    ksteps = max(2, min(args.steps, 5))
    s.block.timing(True)
    for _ in range(ksteps):
        info = s.advance(dt)
        dt = s.next_dt(info.inv_dt_hyp, cfl, 1.1, dt)
    rep = s.block.timing_report()
    s.block.timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if info.nan_events:
        print(f"bench.py: rank {rank}: state is not finite", file=sys.stderr)
        sys.exit(3)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = zones_total * args.steps / (ms_max * 1e-3)

    # ---- end to end through the AdvanceStep contract on host arrays ----------
    e2e = None
    if not args.no_e2e:
        bufs = s.block.data_buffers(pinned=True)
        s.block.download_data(*bufs)
        h2d = sum(b.nbytes for b in bufs if b is not None)
        k2 = max(3, min(args.steps, 10))
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
        e2e = {"value": zones_total * k2 / float(tw.item()), "unit": "zone-updates/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d, "steps": k2,
               "api": "pluto_gpu_advance_data (AdvanceStep on host Data arrays, pinned)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + whole-step figures ----------------
    hbm_peak, peak_src = peaks()
    top = max((k for k in rep if rep[k][1] > 0), key=lambda k: rep[k][0])
    top_ms, top_cnt = rep[top]
    kern_bytes = SWEEP_BYTES_3D.get(top, 0) if dims == 3 else 0
    if dims == 2:
        kern_bytes = SWEEP_BYTES_2D.get(top, 0)
    if args.time_stepping == "hancock":
        kern_bytes = CTU_SWEEP_BYTES_3D.get(top, 0) if dims == 3 else 0
    achieved = kern_bytes * zones_local / (top_ms / top_cnt * 1e-3) / 1e9 if kern_bytes else None
    algo = ALGO.get((solver, recon, dims), dict(bytes=440.0, flops=3300.0))
    step_s = ms_max * 1e-3 / args.steps
    step_gbs = algo["bytes"] * zones_local / step_s / 1e9
    step_tf = algo["flops"] * zones_local / step_s / 1e12
    import ctypes
    from pluto_b200 import load_library
    tf = ctypes.c_double(0.0)
    fp64_measured = None
    if load_library().pluto_gpu_measure_fp64(local, ctypes.byref(tf)) == 0 and tf.value > 0:
        fp64_measured = tf.value
    bound_s = max(algo["bytes"] / (hbm_peak * 1e9), algo["flops"] / (FP64_PEAK_TFLOPS_NOMINAL * 1e12))
    traffic = None
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(tp):      # dram bytes per launch from the committed ncu --set full capture
        tj = json.load(open(tp))
        traffic = tj.get(args.workload, {}).get(top)
    step_kernel_ms = sum(v[0] for v in rep.values()) / ksteps
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": (achieved / hbm_peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": top_ms / top_cnt, "kernel_share_of_step": (top_ms / ksteps) / step_kernel_ms,
                "algorithmic_bytes_per_zone": kern_bytes,
                "note": "the sweeps are bound by the FP64 pipe and instruction issue, not HBM: see roofline_fp64, "
                        "step_roofline and profiles/"}
    # the same kernel against the FP64 pipe, the unit that actually binds it: algorithmic flops per
    # zone and launch from the SURVEY.md 8(d) hand count (PLM 115 + HLLD 360 + RHS 19 + update 8 + C_dt 4
    # = 506 per direction and stage in 3-D; 370 in 2-D; the fused x1+x2 kernel does two directions)
    ndir = 2 if top == "sweep_x1x2" else (1 if top.startswith("sweep") else 0)
    kern_flops = ndir * (506.0 if dims == 3 else 370.0) if (solver, recon) == ("hlld", "plm") else None
    if args.time_stepping == "hancock":
        kern_flops = None          # no SURVEY 8(d) contract figure for the CTU sweeps
    fp64_peak = fp64_measured or FP64_PEAK_TFLOPS_NOMINAL
    roofline_fp64 = None
    if kern_flops:
        ach_tf = kern_flops * zones_local / (top_ms / top_cnt * 1e-3) / 1e12
        roofline_fp64 = {"bound": "fp64", "kernel": top, "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach_tf / fp64_peak, "algorithmic_flops_per_zone": kern_flops,
                         "peak_source": "measured DFMA chain (pluto_gpu_measure_fp64)" if fp64_measured else "nominal"}
    step_roofline = {"algorithmic_bytes_per_zone_update": algo["bytes"], "flops_per_zone_update": algo["flops"],
                     "hbm_gbs": step_gbs, "hbm_frac": step_gbs / hbm_peak,
                     "fp64_tflops": step_tf, "fp64_peak_tflops_nominal": FP64_PEAK_TFLOPS_NOMINAL,
                     "fp64_frac": step_tf / FP64_PEAK_TFLOPS_NOMINAL,
                     "fp64_peak_tflops_measured_dfma_chain": fp64_measured,
                     "fp64_frac_of_measured": (step_tf / fp64_measured) if fp64_measured else None,
                     "stencil_roofline_zone_updates_per_sec_per_gpu": 1.0 / bound_s,
                     "stencil_roofline_frac": (value / world) * bound_s}
    kernels = {k: {"ms_per_step": v[0] / ksteps, "launches_per_step": v[1] / ksteps} for k, v in rep.items() if v[1]}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        n_sample, steps_per = sample_plan(dims)
        copies = os.cpu_count() or 1
        v, wall, kind = cpu_reference_run(problem, dims, recon, solver, cfl, first_dt, n_sample, steps_per, copies,
                                          args.time_stepping)
        cpu_baseline = {"value": v, "unit": "zone-updates/s", "cores": copies, "kind": kind,
                        "sample": f"{copies} concurrent serial copies, {problem} {dims}-D "
                                  f"{'x'.join(str(q) for q in n_sample[:dims])}, {steps_per} steps each, {wall:.1f} s wall"}

    line = {
        "metric": "zone_updates_per_sec", "value": value, "unit": "zone-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "problem": problem, "zones_per_gpu": list(n[:dims]),
                   "global_zones": list(layout.global_n[:dims]), "rank_grid": list(layout.grid),
                   "scheme": f"{solver}+{recon}+ct_uct_contact+" + ("ctu_hancock" if args.time_stepping == "hancock" else "rk2"),
                   "arith": args.arith,
                   "next_dt": "host" if args.host_dt else "device kernel, same dt sequence (tests/test_gpu_parity.py)",
                   "l2": "inputs larger than L2 (state 1.5 GB/GPU at 256^3 vs 126 MB L2)",
                   "device_bytes_per_gpu": s.block.device_bytes},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roofline, "roofline_fp64": roofline_fp64, "step_roofline": step_roofline, "kernels": kernels,
        "cpu_baseline": cpu_baseline,
    }
    if args.dev_lib:
        line["dev_lib"] = args.dev_lib
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
