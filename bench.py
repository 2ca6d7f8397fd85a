#!/usr/bin/env python
"""
Benchmark of the north-star hot path (BASELINE.json): LowStorageRK54 steps of the
single-field scalar-preheating system (ScalarSector(1), V = phi^2/2, conformal FLRW
background) with the 4th-order finite-difference Laplacian (halo_shape = 2) on a
512^3 fp64 lattice; slab decomposition along x over the GPUs (strong scaling, the
global lattice is fixed).

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --steps K --warmup W    # CPU port of the reference

One "step" = one full RK time step = 5 stages; every stage is the fused
stencil + RHS + update kernel with in-kernel energy sums, followed by the two tiny
kernels that fold the sums and advance the FLRW background on the device.
MLUPS counts site-stage updates: global_sites * 5 * steps / seconds / 1e6.
Prints ONE JSON line (rank 0).
"""

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

NUM_STAGES = 5


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--grid", type=int, default=512)
    p.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    p.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--xchunk", type=int, default=0)
    p.add_argument("--variant", type=int, default=0)
    p.add_argument("--nx", type=int, default=0, help="override the global x extent")
    return p.parse_args()


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------

def workload(grid, dtype, world, rank, scaling, nx=0):
    h = 2
    if scaling == "weak":
        grid_shape = (grid * world, grid, grid)
    else:
        grid_shape = (grid, grid, grid)
    if nx:
        grid_shape = (nx, grid, grid)
    box = (5.0 * grid_shape[0] / grid, 5.0, 5.0)
    dx = tuple(L / N for L, N in zip(box, grid_shape))
    dt = 0.1 * min(dx)
    q, r = divmod(grid_shape[0], world)
    nx = q + (1 if r > rank else 0)
    rank_shape = (nx, grid_shape[1], grid_shape[2])
    return dict(h=h, grid_shape=grid_shape, rank_shape=rank_shape, dx=dx, dt=dt,
                mpl=1.0, f0=0.193, df0=-0.142231, amp=1e-6, seed=49279 * (rank + 1),
                np_dtype=np.float64 if dtype == "f64" else np.float32)


def initial_fields(w):
    rng = np.random.default_rng(w["seed"])
    pad = tuple(n + 2 * w["h"] for n in w["rank_shape"])
    f = (w["f0"] + w["amp"] * rng.standard_normal(pad)).astype(w["np_dtype"])
    d = (w["df0"] + w["amp"] * rng.standard_normal(pad)).astype(w["np_dtype"])
    return f[None], d[None]


def background_energy(w):
    # homogeneous-mode energy: kinetic + potential of (f0, df0) at a = 1
    return 0.5 * w["df0"]**2 + 0.5 * w["f0"]**2


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.err = None

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except ValueError:
                    pass
            hnd = nv.nvmlDeviceGetHandleByIndex(idx)
            smax = nv.nvmlDeviceGetMaxClockInfo(hnd, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(hnd, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(hnd)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
                try:
                    power = nv.nvmlDeviceGetPowerUsage(hnd) / 1000.0
                except Exception:
                    power = None
                self.samples.append((sm, smax, reasons, power))
                time.sleep(0.02)
        except Exception as exc:  # noqa: BLE001
            self.err = repr(exc)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0,
                    "reasons": [f"nvml unavailable: {self.err}"]}
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
        reasons = set()
        for _, _, r, _ in self.samples:
            for name, bit in bits.items():
                if r & bit:
                    reasons.add(name)
        powers = [p for _, _, _, p in self.samples if p is not None]
        return {"sm_mhz": float(np.median([s[0] for s in self.samples])),
                "sm_max_mhz": float(self.samples[0][1]), "samples": len(self.samples),
                "power_w_max": max(powers) if powers else None, "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU port of the reference (oracle/oracle_c.c): cpu_baseline and --impl reference
# ---------------------------------------------------------------------------

def cpu_reference_steps(w, nsteps, warmup, budget_s=200.0):
    """Time full LSRK54 steps of the unfused reference schedule on the host cores."""
    import oracle as O
    from oracle import c_oracle as CO
    native = True
    h = w["h"]
    ww = dict(w)
    N = ww["rank_shape"]
    A, B, _ = O.LSRK_TABLES["LowStorageRK54"]
    pot = CO.make_potential(1, [(0.5, (2,))])

    def setup(ws):
        f, d = initial_fields(ws)
        f = f.astype(np.float64)
        d = d.astype(np.float64)
        kf, kd = np.zeros_like(f), np.zeros_like(f)
        lap = np.zeros((1,) + tuple(ws["rank_shape"]))
        gs = float(np.prod(ws["rank_shape"]))
        CO.share_halos(f[0], h, native)
        CO.laplacian(f[0], lap[0], h, ws["dx"], native)
        en = CO.scalar_energy(f, d, lap, h, 1.0, gs, pot, native)
        adot = float(np.sqrt(8 * np.pi / 3 / ws["mpl"]**2 * en.sum()))
        state = np.array([1.0, adot, adot, 0.0, 0.0])
        return f, d, kf, kd, lap, gs, en, state

    def run(ws, st, n):
        f, d, kf, kd, lap, gs, en, state = st
        t0 = time.perf_counter()
        CO.preheating_steps(n, A, B, f, d, kf, kd, lap, h, ws["dx"], ws["dt"], ws["mpl"], gs,
                            pot, state, en, True, native)
        return time.perf_counter() - t0

    st = setup(ww)
    # OpenMP thread count: all hardware threads is not always fastest for this
    # bandwidth-bound loop (SMT / NUMA); try all, half and a quarter, keep the best
    try:
        lib = CO.load(native)
        nmax = os.cpu_count() or CO.num_threads(native)
        best = None
        for nt in sorted({nmax, max(1, nmax // 2), max(1, nmax // 4)}, reverse=True):
            lib.omp_set_num_threads(nt)
            tt = run(ww, st, 1)
            if best is None or tt < best[0]:
                best = (tt, nt)
        lib.omp_set_num_threads(best[1])
    except Exception:  # noqa: BLE001
        pass
    t1 = run(ww, st, 1)  # warm-up step at the chosen thread count; bounds the run
    sample = f"{N[0]}x{N[1]}x{N[2]} lattice, full LSRK54 steps"
    if t1 * (nsteps + warmup) > budget_s and N[1] > 256:
        ww = workload(256, "f64", 1, 0, "strong")
        N = ww["rank_shape"]
        st = setup(ww)
        t1 = run(ww, st, 1)
        sample = (f"reduced to {N[0]}x{N[1]}x{N[2]} lattice (512^3 would exceed the time "
                  "budget on these host cores), full LSRK54 steps")
    if t1 * (nsteps + warmup) > budget_s:
        nsteps = max(1, int(budget_s / t1) - warmup)
        sample += f"; steps cut to {nsteps}"
    for _ in range(max(0, warmup - 1)):
        run(ww, st, 1)
    total = run(ww, st, nsteps)
    sites = float(np.prod(N))
    mlups = sites * NUM_STAGES * nsteps / total / 1e6
    return dict(mlups=mlups, ms_per_step=total / nsteps * 1e3, steps=nsteps,
                cores=CO.num_threads(native), sample=sample, state=st[7].tolist())


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.grid, "f64", 1, 0, "strong")
    r = cpu_reference_steps(w, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": r["mlups"],
        "unit": "1e6 site-stage updates/s", "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "scalar_preheating single-field phi^2/2, LowStorageRK54, "
                               "4th-order FD (h=2), fp64, unfused reference schedule "
                               "(share_halos -> lap -> stage -> energy -> expansion)",
                   "grid": args.grid},
        "cpu_baseline": {"value": r["mlups"], "unit": "1e6 site-stage updates/s",
                         "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["mlups"], "unit": "1e6 site-stage updates/s",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C/OpenMP port of the reference's schedule (oracle/oracle_c.c): "
                "pyopencl/loopy/POCL are not installable in this image",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------

def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import pystella_b200 as ps
    from pystella_b200 import _kernels

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N"

    w = workload(args.grid, args.dtype, world, rank, args.scaling, args.nx)
    h, dx, dt = w["h"], w["dx"], w["dt"]
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    S = 8 if args.dtype == "f64" else 4
    global_sites = float(np.prod(w["grid_shape"]))
    local_sites = float(np.prod(w["rank_shape"]))

    decomp = ps.DomainDecomposition((world, 1, 1), h, rank_shape=w["rank_shape"])
    sector = ps.ScalarSector(1, potential=lambda f: f[0]**2 / 2)
    tune = {"xchunk": args.xchunk, "variant": args.variant}

    f_host, d_host = initial_fields(w)
    pin = [torch.from_numpy(x).pin_memory() for x in (f_host, d_host)]
    e0 = background_energy(w)

    def make_state():
        stepper = ps.LowStorageRK54(sector, halo_shape=h, dt=dt, dx=dx, decomp=decomp,
                                    MapKernel=ps.FusedStencilStage, energy=True, tune=tune)
        f = pin[0].to(dev, non_blocking=True)
        d = pin[1].to(dev, non_blocking=True)
        expand = ps.DeviceExpansion(e0, ps.LowStorageRK54, mpl=w["mpl"],
                                    grid_size=global_sites, decomp=decomp, nscalars=1)
        return stepper, f, d, expand

    # ---------------- device-resident loop -> `value` ------------------------
    stepper, f, dfdt, expand = make_state()
    launches_per_step = [0]

    def one_step():
        for s in range(NUM_STAGES):
            stepper(s, f=f, dfdt=dfdt, a=expand.a, hubble=expand.hubble)
            expand.step(s, stepper, dt)
    launches_per_step[0] = NUM_STAGES * 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    # launches per stage: stage kernel(s) + energy fold + expansion; the NCCL fallback
    # splits the stage into boundary + interior launches
    halo_mode = getattr(stepper, "_halo_mode", "local")
    launches_per_step[0] = NUM_STAGES * ((2 if halo_mode == "nccl" else 1) + 2)

    graph = None
    if not args.no_graph:
        try:
            # two steps per replay: the ping-pong storage of f returns to its start
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                one_step()
                one_step()
            torch.cuda.current_stream().wait_stream(side)
            barrier()
            with torch.cuda.graph(g):
                one_step()
                one_step()
            graph = g
            graph.replay()
            barrier()
        except Exception as exc:  # noqa: BLE001
            graph = None
            if rank == 0:
                print(f"[bench] CUDA graph capture unavailable ({exc!r}); timing eager "
                      "launches", file=sys.stderr)
            barrier()

    steps = args.steps
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    e0_ev = torch.cuda.Event(enable_timing=True)
    e1_ev = torch.cuda.Event(enable_timing=True)
    e0_ev.record()
    if graph is not None:
        for _ in range(steps // 2):
            graph.replay()
        for _ in range(steps % 2):
            one_step()
    else:
        for _ in range(steps):
            one_step()
    e1_ev.record()
    barrier()
    ms_total = e0_ev.elapsed_time(e1_ev)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    ms_per_step = ms_total / steps
    value = global_sites * NUM_STAGES * steps / (ms_total * 1e-3) / 1e6
    bg = expand.get()
    assert np.isfinite(bg["a"]) and bg["a"] > 1.0, "background did not evolve"

    # ---------------- roofline of the dominant kernel (eager, per-launch events) ----
    if os.environ.get("PSB_TIMELINE", "0") == "1":
        from pystella_b200 import fused as _fused
        _fused._timeline.clear()
    stage_ms = []
    reps = max(2, min(steps, 6))
    for _ in range(reps):
        for s in range(NUM_STAGES):
            a_ev = torch.cuda.Event(enable_timing=True)
            b_ev = torch.cuda.Event(enable_timing=True)
            a_ev.record()
            stepper(s, f=f, dfdt=dfdt, a=expand.a, hubble=expand.hubble)
            b_ev.record()
            expand.step(s, stepper, dt)
            stage_ms.append((s, a_ev, b_ev))
    torch.cuda.synchronize()
    if os.environ.get("PSB_TIMELINE", "0") == "1" and rank == 0:
        from pystella_b200.fused import timeline_report
        print("[timeline ms since stage begin]", timeline_report(), file=sys.stderr)
    per_stage = {}
    for s, a_ev, b_ev in stage_ms:
        per_stage.setdefault(s, []).append(a_ev.elapsed_time(b_ev))
    # algorithmic bytes per launch: read u,k + write u,k for f and dfdt; A[0] = 0 skips
    # reading the two twins at stage 0 ("bytes per site-stage")
    alg = {s: local_sites * (6 if s == 0 else 8) * S for s in range(NUM_STAGES)}
    tot_bytes = sum(alg[s] * len(per_stage[s]) for s in per_stage)
    tot_ms = sum(sum(v) for v in per_stage.values())
    achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json hbm_gbs (torch copy, measured on this pool)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "fused_stage_traffic.json")
    if os.path.exists(tpath) and args.grid == 512 and world == 1 and args.dtype == "f64":
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "psb::fused_stage_kernel (per-launch CUDA events, eager launches)",
                "algorithmic_bytes_per_site_stage": {"stage0": 6 * S, "stages1-4": 8 * S},
                "ms_per_launch": {str(s): float(np.mean(v)) for s, v in per_stage.items()}}

    # ---------------- end-to-end through the public API with host buffers ---------
    e2e = None
    e2e_dropin = None
    if not args.no_e2e:
        del stepper, f, dfdt, expand
        torch.cuda.empty_cache()
        e2e = run_e2e_resident(args, ps, torch, dist, w, decomp, sector, pin, dev, world,
                               rank, global_sites, tune, e0)
        torch.cuda.empty_cache()
        e2e_dropin = run_e2e(args, ps, torch, dist, w, decomp, sector, pin, dev, world,
                             rank, global_sites, tune)

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = cpu_reference_steps(workload(args.grid, "f64", 1, 0, "strong"), 2, 1,
                                    budget_s=40.0)
            cpu = {"value": r["mlups"], "unit": "1e6 site-stage updates/s",
                   "cores": r["cores"], "kind": "port",
                   "sample": r["sample"] + f" ({r['steps']} timed after 1 warm-up), "
                   "unfused reference schedule, oracle/oracle_c.c"}
        except Exception as exc:  # noqa: BLE001
            cpu = {"value": None, "unit": "1e6 site-stage updates/s", "cores": None,
                   "kind": "port", "sample": f"failed: {exc!r}"}

    if rank == 0:
        line = {
            "metric": "MLUPS", "value": value, "unit": "1e6 site-stage updates/s",
            "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": "scalar_preheating single-field phi^2/2 + FLRW background, "
                            "LowStorageRK54, 4th-order FD (h=2), "
                            f"{'x'.join(str(n) for n in w['grid_shape'])} {args.dtype}",
                "grid_shape": list(w["grid_shape"]), "rank_shape": list(w["rank_shape"]),
                "parallelism": f"x-slabs over {world} GPU(s)",
                "halo_exchange": {"local": "periodic wrap written by the stage kernel",
                                  "peer": "stage kernel stores boundary planes into the "
                                          "neighbours' halos over NVLink (peer memory)",
                                  "nccl": "NCCL send/recv on a side stream"}[halo_mode],
                "l2": "inputs larger than L2 (4 field arrays of "
                      f"{local_sites * S / 1e6:.0f} MB each per GPU vs 126 MB L2)",
                "launch": "CUDA graph (2 steps per replay)" if graph is not None
                          else "eager launches",
                "full_step_mlups": value / NUM_STAGES,
            },
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "e2e_dropin": e2e_dropin,
            "gpu_launches": launches_per_step[0] * steps,
            "background": {"a": bg["a"], "hubble": bg["hubble"]},
        }
        print(json.dumps(line), flush=True)
    # leave without tearing NCCL down collectively (captured graphs keep the
    # communicator busy; a hang here would only waste the lease)
    sys.stdout.flush()
    if world > 1:
        try:
            dist.barrier()
        except Exception:  # noqa: BLE001
            pass
        os._exit(0)


def run_e2e_resident(args, ps, torch, dist, w, decomp, sector, pin, dev, world, rank,
                     global_sites, tune, e0):
    """
    End to end through the public API as INTEGRATION.md recommends it: initial
    conditions uploaded from pinned host memory, the time loop of fused stages
    (in-kernel energy sums) + DeviceExpansion called stage by stage from Python (no
    CUDA graph), after EVERY step the background and energies are read back to the
    host (what a run logs per step), final fields downloaded to pinned host memory.
    Wall clock over the whole job, max over ranks.
    """
    h, dx, dt = w["h"], w["dx"], w["dt"]
    steps = args.steps
    stepper = ps.LowStorageRK54(sector, halo_shape=h, dt=dt, dx=dx, decomp=decomp,
                                MapKernel=ps.FusedStencilStage, energy=True, tune=tune)
    out_host = [torch.empty_like(p).pin_memory() for p in pin]

    def job(nsteps):
        f = pin[0].to(dev, non_blocking=True)
        dfdt = pin[1].to(dev, non_blocking=True)
        stepper.reset()
        expand = ps.DeviceExpansion(e0, ps.LowStorageRK54, mpl=w["mpl"],
                                    grid_size=global_sites, decomp=decomp, nscalars=1)
        log, d2h = [], 0
        for _ in range(nsteps):
            for s in range(NUM_STAGES):
                stepper(s, f=f, dfdt=dfdt, a=expand.a, hubble=expand.hubble)
                expand.step(s, stepper, dt)
            rec = expand.get()          # device -> host: a, adot, hubble + 5 energies
            log.append((rec["a"], rec["total"], rec["pressure"]))
            d2h += (5 + 5) * 8
        out_host[0].copy_(f, non_blocking=True)
        out_host[1].copy_(dfdt, non_blocking=True)
        torch.cuda.synchronize()
        return log, d2h

    job(1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    log, d2h_scalars = job(steps)
    if world > 1:
        dist.barrier()
    sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    field_bytes = sum(p.numel() * p.element_size() for p in pin)
    return {"value": global_sites * NUM_STAGES * steps / sec / 1e6,
            "unit": "1e6 site-stage updates/s", "steps": steps,
            "h2d_bytes_per_step": field_bytes / steps,
            "d2h_bytes_per_step": field_bytes / steps + d2h_scalars / steps,
            "what": "public API, device-resident loop (INTEGRATION.md): upload ICs from "
                    "pinned host memory, Python-driven fused stages + DeviceExpansion, "
                    "background and energies read back to the host after every step, final "
                    "fields downloaded; wall clock incl. all copies, upload/download "
                    f"amortised over {steps} steps",
            "a_final": float(log[-1][0])}


def run_e2e(args, ps, torch, dist, w, decomp, sector, pin, dev, world, rank,
            global_sites, tune):
    """
    The loop a pystella user writes (examples/scalar_preheating.py:258-266) through
    the public API: fields uploaded from pinned host memory, per stage the stepper,
    then FiniteDifferencer + Reduction (device -> host energies) feeding the HOST
    Expansion, whose scalars go back into the next stage; final fields downloaded.
    """
    h, dx, dt = w["h"], w["dx"], w["dt"]
    steps = max(1, min(args.steps, 10))
    stepper = ps.LowStorageRK54(sector, halo_shape=h, dt=dt, dx=dx, decomp=decomp,
                                MapKernel=ps.FusedStencilStage, tune=tune)
    derivs = ps.FiniteDifferencer(decomp, h, dx, rank_shape=w["rank_shape"])
    reduce_energy = ps.Reduction(decomp, sector, halo_shape=h, callback=ps.get_rho_and_p,
                                 rank_shape=w["rank_shape"], grid_size=global_sites)
    tdtype = pin[0].dtype
    lap_f = torch.zeros((1,) + tuple(w["rank_shape"]), dtype=tdtype, device=dev)
    out_host = [torch.empty_like(p).pin_memory() for p in pin]

    def compute_energy(f, dfdt, a):
        derivs(None, fx=f, lap=lap_f)
        return reduce_energy(None, f=f, dfdt=dfdt, lap_f=lap_f, a=np.array(a))

    def job(nsteps):
        f = pin[0].to(dev, non_blocking=True)
        dfdt = pin[1].to(dev, non_blocking=True)
        stepper.reset()
        energy = compute_energy(f, dfdt, 1.)
        expand = ps.Expansion(energy["total"], ps.LowStorageRK54, mpl=w["mpl"])
        d2h = 0
        for _ in range(nsteps):
            for s in range(NUM_STAGES):
                stepper(s, f=f, dfdt=dfdt, a=expand.a, hubble=expand.hubble)
                expand.step(s, energy["total"], energy["pressure"], dt)
                energy = compute_energy(f, dfdt, expand.a)
                d2h += 3 * 8
        out_host[0].copy_(f, non_blocking=True)
        out_host[1].copy_(dfdt, non_blocking=True)
        torch.cuda.synchronize()
        return expand, d2h

    job(1)  # warm-up (allocations, first-call paths)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    expand, d2h_scalars = job(steps)
    if world > 1:
        dist.barrier()
    sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    field_bytes = sum(p.numel() * p.element_size() for p in pin)
    return {"value": global_sites * NUM_STAGES * steps / sec / 1e6,
            "unit": "1e6 site-stage updates/s", "steps": steps,
            "h2d_bytes_per_step": field_bytes / steps,
            "d2h_bytes_per_step": field_bytes / steps + d2h_scalars / steps,
            "what": "reference loop shape unchanged (examples/scalar_preheating.py:258-266): "
                    "upload ICs from pinned host memory, per stage fused stepper + "
                    "FiniteDifferencer + Reduction (energies read back to the host) + host "
                    "Expansion, download final fields; wall clock incl. all copies, "
                    "upload/download amortised "
                    f"over {steps} steps",
            "a_final": float(expand.a[0])}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)
