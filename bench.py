#!/usr/bin/env python
"""
Benchmark of the north-star hot path (BASELINE.json): LowStorageRK54 steps of the
scalar-preheating system with the 4th-order finite-difference Laplacian (halo_shape = 2),
every RK stage ONE fused stencil + RHS + update kernel (in-kernel energy sums; its last CTA
folds them, runs the inter-GPU exchange and advances the FLRW background), x-slab
decomposition over the GPUs.

    python bench.py --gpus N --steps K --warmup W [--config c2|c3|c4|c5]   # this framework
    python bench.py --impl reference --steps K --warmup W [--config ...]    # CPU port

Configurations (BASELINE.json `configs`):
    c3 (default)  single field phi^2/2, 512^3 fp64, strong scaling   <- the headline metric
    c2            the same at 256^3 on one GPU
    c4            two fields + gravitational-wave tensor sector, 512^3 fp32, strong scaling
    c5            single field, fp32, weak scaling: (128, 1024, 1024) per GPU (1024^3 at N=8)

One "step" = one full RK time step = 5 stages.  MLUPS counts site-stage updates:
global_sites * 5 * steps / seconds / 1e6.  Prints ONE JSON line (rank 0).  After the timed
region the run restarts from the seeded initial conditions, integrates 2 steps on all ranks
through the public API, gathers an x-window that straddles a rank boundary (the periodic
boundary at N = 1) and compares it with the CPU oracle on rank 0: `"parity"` in the line.

The default run (N = 1, config c3, no overrides) afterwards collects `"extras"`: the other
single-GPU configurations (c2, c5, c4) and the diagnostics kernels (`--diagnostics`:
histograms, spectrum binning, stencil map), each in its own subprocess of this script under a
time limit (`--extras-budget`, `--no-extras`); the headline numbers are measured, and echoed to
stderr, before any of them starts.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_STAGES = 5
UNIT = "1e6 site-stage updates/s"

CONFIGS = {
    "c2": dict(grid=(256, 256, 256), dtype="f64", scaling="strong", system="scalar1",
               amp=1e-6),
    "c3": dict(grid=(512, 512, 512), dtype="f64", scaling="strong", system="scalar1",
               amp=1e-6),
    # amp 1e-2: fp32 must resolve the gradients of the fluctuations that source hij
    "c4": dict(grid=(512, 512, 512), dtype="f32", scaling="strong", system="scalar2gw",
               amp=1e-2),
    "c5": dict(grid=(128, 1024, 1024), dtype="f32", scaling="weak", system="scalar1",
               amp=1e-6),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    p.add_argument("--grid", type=int, default=0, help="override: cubic grid of this size")
    p.add_argument("--dtype", default=None, choices=["f64", "f32"])
    p.add_argument("--scaling", default=None, choices=["strong", "weak"])
    p.add_argument("--nx", type=int, default=0, help="override the global x extent")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-parity", action="store_true")
    p.add_argument("--xchunk", type=int, default=0)
    p.add_argument("--variant", type=int, default=0)
    p.add_argument("--no-extras", action="store_true",
                   help="skip the extra records of the default N=1 run (other configurations "
                        "and the diagnostics kernels, each in its own bounded subprocess)")
    p.add_argument("--extras-budget", type=float, default=200.0,
                   help="wall-clock budget (s) of all extra records together")
    p.add_argument("--diagnostics", action="store_true",
                   help="time the diagnostics kernels only (histograms, spectrum binning) "
                        "and print their JSON record")
    p.add_argument("--pdl", type=int, default=-1, choices=[-1, 0, 1],
                   help="programmatic dependent launch of the stage kernels (-1: library default)")
    return p.parse_args()


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------

def resolve_config(args):
    cfg = dict(CONFIGS[args.config])
    cfg["name"] = args.config
    if args.grid:
        cfg["grid"] = (args.grid,) * 3
    if args.dtype:
        cfg["dtype"] = args.dtype
    if args.scaling:
        cfg["scaling"] = args.scaling
    if args.nx:
        cfg["grid"] = (args.nx,) + tuple(cfg["grid"][1:])
    return cfg


def workload(cfg, world, rank):
    h = 2
    g = cfg["grid"]
    grid_shape = (g[0] * world, g[1], g[2]) if cfg["scaling"] == "weak" else tuple(g)
    dxs = 5.0 / g[1]
    dx = (dxs, dxs, dxs)
    dt = 0.1 * dxs
    q, r = divmod(grid_shape[0], world)
    nx = q + (1 if r > rank else 0)
    start = rank * q + min(rank, r)
    ns = 2 if cfg["system"] == "scalar2gw" else 1
    return dict(h=h, grid_shape=grid_shape, rank_shape=(nx, g[1], g[2]), x_start=start, dx=dx,
                dt=dt, mpl=1.0, f0=[0.193, 0.0][:ns], df0=[-0.142231, 0.0][:ns],
                amp=cfg["amp"], seed=49279 * (rank + 1), nscalars=ns,
                gw=cfg["system"] == "scalar2gw", dtype=cfg["dtype"],
                np_dtype=np.float64 if cfg["dtype"] == "f64" else np.float32)


def potential_terms(w):
    """V as (coef, powers) monomials: phi^2/2 or the example's two-field potential
    (examples/scalar_preheating.py:98-105 with mphi=1.2e-6, gsq=2.5e-7, in units of mphi)."""
    if w["nscalars"] == 1:
        return [(0.5, (2,))]
    return [(0.5, (2, 0)), (2.5e-7 / 2 / 1.2e-6**2, (2, 2))]


def potential_callable(w):
    terms = potential_terms(w)

    def V(f):
        out = 0
        for c, pw in terms:
            t = c
            for m, e in enumerate(pw):
                if e:
                    t = t * f[m]**e
            out = out + t
        return out
    return V


def initial_fields(w):
    """Padded (nscalars, Nx+2h, Ny+2h, Nz+2h) host arrays of this rank (seeded per rank)."""
    rng = np.random.default_rng(w["seed"])
    pad = tuple(n + 2 * w["h"] for n in w["rank_shape"])
    f = np.stack([(f0 + w["amp"] * rng.standard_normal(pad)) for f0 in w["f0"]])
    d = np.stack([(d0 + w["amp"] * rng.standard_normal(pad)) for d0 in w["df0"]])
    return f.astype(w["np_dtype"]), d.astype(w["np_dtype"])


def background_energy(w):
    # homogeneous-mode energy: kinetic + potential of (f0, df0) at a = 1
    e = sum(0.5 * d**2 for d in w["df0"])
    for c, pw in potential_terms(w):
        t = c
        for m, p in enumerate(pw):
            t *= w["f0"][m]**p
        e += t
    return e


def config_dict(cfg, w, world):
    """The `config` object: identical in both arms (workload only, no run-time details)."""
    sysname = {"scalar1": "scalar_preheating single-field phi^2/2 + FLRW background",
               "scalar2gw": "two-field preheating (phi^2/2 + g^2 phi^2 chi^2/2) + "
                            "gravitational-wave tensor sector (6 hij + 6 dhijdt) + FLRW "
                            "background"}[cfg["system"]]
    S = 8 if cfg["dtype"] == "f64" else 4
    narr = 4 * w["nscalars"] + (24 if w["gw"] else 0)
    per = float(np.prod(w["grid_shape"])) / world * S / 1e6
    return {
        "workload": f"{sysname}, LowStorageRK54, 4th-order FD (h=2), "
                    f"{'x'.join(str(n) for n in w['grid_shape'])} {cfg['dtype']}",
        "name": cfg["name"], "grid_shape": list(w["grid_shape"]), "dtype": cfg["dtype"],
        "scaling": cfg["scaling"], "stepper": "LowStorageRK54", "halo_shape": 2,
        "parallelism": f"x-slabs over {world} GPU(s)",
        "l2": f"inputs larger than L2 ({narr} field arrays of {per:.0f} MB each per GPU vs "
              "126 MB L2)",
    }


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

def _cpu_threads_policy():
    """Thread policy of the timed CPU baseline: threads spread over the cores the process may
    use (OMP_PROC_BIND=spread), pages first-touched in parallel by the same threads.  The
    thread COUNT is calibrated (see _calibrate_threads): this bandwidth-bound loop is not
    fastest with every hardware thread (measured on one 128-thread host: 172 MLUPS with 128
    threads, 600-1200 with 32)."""
    os.environ.setdefault("OMP_PROC_BIND", "spread")
    os.environ.setdefault("OMP_PLACES", "threads")
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _calibrate_threads(lib, nmax, setup, run, cfg):
    """Best OpenMP thread count among nmax, nmax/2, nmax/4, nmax/8 on a 128^3 sample of the
    same system (first touch repeated for every candidate); a few tenths of a second each."""
    small = dict(cfg, scaling="strong", grid=(128, 128, 128))
    best = None
    for nt in sorted({nmax, max(1, nmax // 2), max(1, nmax // 4), max(1, nmax // 8)},
                     reverse=True):
        lib.omp_set_num_threads(nt)
        st = setup(small)
        run(st, 1)
        t = min(run(st, 1) for _ in range(2))
        if best is None or t < best[0]:
            best = (t, nt)
        del st
    return best[1]


def cpu_reference_steps(cfg, nsteps, warmup, budget_s=150.0):
    """Time full LSRK54 steps of the unfused reference schedule on the host cores: a bounded
    sample of the configuration's workload (the lattice is shrunk until one step fits the
    budget), best of 3 repetitions of `nsteps` steps."""
    nthreads = _cpu_threads_policy()
    import oracle as O
    from oracle import c_oracle as CO
    native = True
    lib = CO.load(native)
    A, B, _ = O.LSRK_TABLES["LowStorageRK54"]

    def setup(cfgx):
        w = workload(cfgx, 1, 0)
        h = w["h"]
        ns = w["nscalars"]
        pot = CO.make_potential(ns, potential_terms(w))
        f0, d0 = initial_fields(w)
        N = w["rank_shape"]
        f, d = CO.first_touch_copy(f0, native), CO.first_touch_copy(d0, native)
        del f0, d0
        kf, kd = CO.first_touch_zeros(f.shape, native), CO.first_touch_zeros(f.shape, native)
        gs = float(np.prod(N))
        st = dict(w=w, pot=pot, f=f, d=d, kf=kf, kd=kd, gs=gs)
        if w["gw"]:
            pad = f.shape[1:]
            for name in ("hij", "dhij", "kh", "kdh"):
                st[name] = CO.first_touch_zeros((6,) + pad, native)
            lap_f, dfdx, lap_hij = (CO.first_touch_zeros(s, native)
                                    for s in ((ns,) + N, (ns, 3) + N, (6,) + N))
            for n in range(ns):
                CO.share_halos(f[n], h, native)
                CO.laplacian(f[n], lap_f[n], h, w["dx"], native)
                CO.gradient(f[n], dfdx[n], h, w["dx"], native)
            st.update(lap=lap_f, dfdx=dfdx, lap_hij=lap_hij)
        else:
            lap = CO.first_touch_zeros((ns,) + N, native)
            for n in range(ns):
                CO.share_halos(f[n], h, native)
                CO.laplacian(f[n], lap[n], h, w["dx"], native)
            st["lap"] = lap
        en = CO.scalar_energy(f, d, st["lap"], h, 1.0, gs, pot, native)
        adot = float(np.sqrt(8 * np.pi / 3 / w["mpl"]**2 * en.sum()))
        st["en"] = en
        st["state"] = np.array([1.0, adot, adot, 0.0, 0.0])
        return st

    def run(st, n):
        w = st["w"]
        t0 = time.perf_counter()
        if w["gw"]:
            CO.preheating_gw_steps(n, A, B, st["f"], st["d"], st["kf"], st["kd"], st["lap"],
                                   st["dfdx"], st["hij"], st["dhij"], st["kh"], st["kdh"],
                                   st["lap_hij"], w["h"], w["dx"], w["dt"], w["mpl"], st["gs"],
                                   st["pot"], st["state"], st["en"], True, native)
        else:
            CO.preheating_steps(n, A, B, st["f"], st["d"], st["kf"], st["kd"], st["lap"],
                                w["h"], w["dx"], w["dt"], w["mpl"], st["gs"], st["pot"],
                                st["state"], st["en"], True, native)
        return time.perf_counter() - t0

    # the sample: shrink the lattice (x first, then all axes) until a step fits the budget;
    # the tensor configuration needs 50 fp64 arrays, so it starts at 256^3
    try:
        nthreads = _calibrate_threads(lib, nthreads, setup, run, cfg)
        lib.omp_set_num_threads(nthreads)
    except Exception:  # noqa: BLE001
        pass
    cfgx = dict(cfg, scaling="strong")
    g = list(workload(cfg, 1, 0)["grid_shape"])
    if cfg["system"] == "scalar2gw":
        g = [min(n, 256) for n in g]
    cfgx["grid"] = tuple(g)
    reduced = tuple(g) != tuple(workload(cfg, 1, 0)["grid_shape"])
    st = setup(cfgx)
    t1 = run(st, 1)
    reps = 3
    while t1 * (reps * nsteps + warmup + 1) > budget_s:
        if reps > 1:
            reps = 1
            continue
        if min(cfgx["grid"]) <= 64:
            break
        cfgx["grid"] = tuple(max(64, n // 2) for n in cfgx["grid"])
        reduced = True
        del st
        st = setup(cfgx)
        t1 = run(st, 1)
    for _ in range(max(0, warmup - 1)):
        run(st, 1)
    times = [run(st, nsteps) for _ in range(reps)]
    total = min(times)
    N = st["w"]["rank_shape"]
    sites = float(np.prod(N))
    mlups = sites * NUM_STAGES * nsteps / total / 1e6
    sample = (f"{N[0]}x{N[1]}x{N[2]} lattice" + (" (reduced from the configuration's "
              "lattice to bound the CPU time and memory)" if reduced else "")
              + f", full LSRK54 steps in fp64, best of {reps} x {nsteps} steps after "
              f"{max(warmup, 1)} warm-up, {CO.num_threads(native)} OpenMP threads (count "
              "calibrated on a 128^3 sample, OMP_PROC_BIND=spread, parallel first touch), "
              "unfused reference schedule "
              "(share_halos -> lap -> stage -> energy -> expansion), oracle/oracle_c.c")
    return dict(mlups=mlups, ms_per_step=total / nsteps * 1e3, steps=nsteps,
                cores=CO.num_threads(native), sample=sample, spread=[min(times), max(times)])


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = resolve_config(args)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    r = cpu_reference_steps(cfg, max(1, args.steps), max(1, args.warmup), budget_s=240.0)
    w = workload(cfg, world, 0)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": r["mlups"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": max(1, args.warmup),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(cfg, w, world),
        "cpu_baseline": {"value": r["mlups"], "unit": UNIT, "cores": r["cores"],
                         "kind": "port", "sample": r["sample"],
                         "seconds_min_max": r["spread"]},
        "e2e": {"value": r["mlups"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C/OpenMP port of the reference's schedule (oracle/oracle_c.c): "
                "pyopencl/loopy/POCL are not installable in this image; the CPU arithmetic "
                "is fp64 for every configuration",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------

class Job:
    """The configuration's system through the public API (what a user writes)."""

    def __init__(self, ps, torch, cfg, w, decomp, dev, tune, global_sites):
        self.ps, self.torch, self.w, self.dev = ps, torch, w, dev
        self.decomp, self.tune, self.global_sites = decomp, tune, global_sites
        self.scalar = ps.ScalarSector(w["nscalars"], potential=potential_callable(w))
        self.sectors = [self.scalar]
        if w["gw"]:
            self.sectors.append(ps.TensorPerturbationSector([self.scalar]))
        self.e0 = background_energy(w)
        f_host, d_host = initial_fields(w)
        self.pin = [torch.from_numpy(x).pin_memory() for x in (f_host, d_host)]
        self.tdtype = self.pin[0].dtype

    def make_stepper(self, energy=True):
        w = self.w
        sec = self.sectors if len(self.sectors) > 1 else self.sectors[0]
        return self.ps.LowStorageRK54(sec, halo_shape=w["h"], dt=w["dt"], dx=w["dx"],
                                      decomp=self.decomp, MapKernel=self.ps.FusedStencilStage,
                                      energy=energy, tune=self.tune)

    def upload(self):
        torch = self.torch
        arrays = dict(f=self.pin[0].to(self.dev, non_blocking=True),
                      dfdt=self.pin[1].to(self.dev, non_blocking=True))
        if self.w["gw"]:
            pad = tuple(self.pin[0].shape[1:])
            arrays["hij"] = torch.zeros((6,) + pad, dtype=self.tdtype, device=self.dev)
            arrays["dhijdt"] = torch.zeros((6,) + pad, dtype=self.tdtype, device=self.dev)
        return arrays

    def make_expansion(self):
        return self.ps.DeviceExpansion(self.e0, self.ps.LowStorageRK54, mpl=self.w["mpl"],
                                       grid_size=self.global_sites, decomp=self.decomp,
                                       nscalars=self.w["nscalars"])

    def field_bytes(self):
        return sum(p.numel() * p.element_size() for p in self.pin)


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import pystella_b200 as ps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N"

    cfg = resolve_config(args)
    w = workload(cfg, world, rank)
    h, dt = w["h"], w["dt"]
    S = 8 if cfg["dtype"] == "f64" else 4
    global_sites = float(np.prod(w["grid_shape"]))
    local_sites = float(np.prod(w["rank_shape"]))
    ns = w["nscalars"]

    decomp = ps.DomainDecomposition((world, 1, 1), h, rank_shape=w["rank_shape"])
    tune = {"xchunk": args.xchunk, "variant": args.variant}
    if args.pdl >= 0:
        tune["pdl"] = bool(args.pdl)
    job = Job(ps, torch, cfg, w, decomp, dev, tune, global_sites)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident loop -> `value` ------------------------
    stepper = job.make_stepper()
    arrays = job.upload()
    expand = job.make_expansion()
    expand.attach(stepper)

    def one_step():
        for s in range(NUM_STAGES):
            stepper(s, a=expand.a, hubble=expand.hubble, **arrays)
            expand.step(s, stepper, dt)

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    halo_mode = getattr(stepper, "_halo_mode", "local")
    # kernels per stage: the scalar stage (+ the tensor stage); the NCCL fallback splits the
    # scalar stage into boundary + interior launches and keeps the fold / expansion kernels
    kernels_per_stage = {"local": 1 + (1 if w["gw"] else 0),
                         # + the inter-GPU exchange / background kernel
                         "peer": 2 + (1 if w["gw"] else 0), "nccl": 4}[halo_mode]
    launches_per_step = NUM_STAGES * kernels_per_stage

    graph = None
    if not args.no_graph:
        try:
            # two steps per replay: the ping-pong storage of f returns to its start
            graph = ps.capture_steps(one_step, steps_per_replay=2, warmup=1,
                                     num_stages=NUM_STAGES)
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
    if steps % 2 and graph is not None:
        one_step()      # back to an even stage count before anything else is replayed

    # ---------------- roofline of the dominant kernel (eager, per-launch events) ----
    stepper._timing_hook = []
    reps = max(2, min(steps, 6))
    for _ in range(reps):
        # keep the device busy while the host queues the five stages, so that every event
        # pair brackets exactly one kernel (no idle time of a slow host inside the interval)
        torch.cuda._sleep(20_000_000)
        one_step()
    torch.cuda.synchronize()
    hook, stepper._timing_hook = stepper._timing_hook, None
    per = {}   # (system index, stage) -> [ms]
    for stage, idx, a_ev, b_ev in hook:
        per.setdefault((idx, stage), []).append(a_ev.elapsed_time(b_ev))
    # algorithmic bytes per launch: read u,k + write u,k per evolved array; A[0] = 0 skips
    # reading the twins at stage 0 ("bytes per site-stage")
    dom = 1 if w["gw"] else 0            # dominant kernel: tensor stage (c4) / scalar stage
    narr = 6 if w["gw"] else ns          # stencilled unknowns handled by that kernel
    alg = {s: local_sites * narr * (6 if s == 0 else 8) * S for s in range(NUM_STAGES)}
    tot_bytes = sum(alg[s] * len(per[(dom, s)]) for s in range(NUM_STAGES))
    tot_ms = sum(sum(per[(dom, s)]) for s in range(NUM_STAGES))
    achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json hbm_gbs (torch copy, measured on this pool)"
    else:
        peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "stage_traffic_r2.json")
    if os.path.exists(tpath) and world == 1:
        rec = json.load(open(tpath)).get(cfg["name"])
        if rec and list(rec.get("grid_shape", [])) == list(w["grid_shape"]):
            traffic, traffic_src = rec.get("dram_bytes_per_launch"), rec.get("source")
    kernel = ("psb::tensor_stage_ws_kernel" if w["gw"] else "psb::fused_stage_ws_kernel")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "kernel": kernel + " (per-launch CUDA events on the launching stream, eager "
                                   "launches, after the timed region)",
                "algorithmic_bytes_per_site_stage": {"stage0": narr * 6 * S,
                                                     "stages1-4": narr * 8 * S},
                "ms_per_launch": {str(s): float(np.mean(per[(dom, s)]))
                                  for s in range(NUM_STAGES)},
                "frac_per_stage": {str(s): alg[s] / (float(np.mean(per[(dom, s)])) * 1e-3)
                                   / 1e9 / peak for s in range(NUM_STAGES)}}
    if w["gw"]:
        roofline["scalar_stage_ms_per_launch"] = {
            str(s): float(np.mean(per[(0, s)])) for s in range(NUM_STAGES)}
        whole = sum(float(np.mean(per[(0, s)])) + float(np.mean(per[(1, s)]))
                    for s in range(NUM_STAGES))
        whole_bytes = sum(local_sites * (ns + 6) * (6 if s == 0 else 8) * S
                          for s in range(NUM_STAGES))
        roofline["whole_stage"] = {
            "algorithmic_bytes_per_site_stage": (ns + 6) * 8 * S,
            "achieved": whole_bytes / (whole * 1e-3) / 1e9,
            "frac": whole_bytes / (whole * 1e-3) / 1e9 / peak}

    # ---------------- parity of the path that was just timed ------------------------
    parity = None
    if not args.no_parity:
        del stepper, arrays, expand, graph
        torch.cuda.empty_cache()
        try:
            parity = run_parity(job, cfg, w, world, rank, dist)
        except Exception as exc:  # noqa: BLE001
            parity = {"ok": False, "error": repr(exc)}
        torch.cuda.empty_cache()
    else:
        del stepper, arrays, expand, graph
        torch.cuda.empty_cache()

    # ---------------- end-to-end through the public API with host buffers ---------
    e2e = None
    e2e_dropin = None
    if not args.no_e2e:
        e2e = run_e2e_resident(args, job, dist, world)
        torch.cuda.empty_cache()
        if not w["gw"]:
            e2e_dropin = run_e2e_dropin(args, job, dist, world)
            torch.cuda.empty_cache()

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = cpu_reference_steps(cfg, 2, 1, budget_s=30.0)
            cpu = {"value": r["mlups"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                   "sample": r["sample"], "seconds_min_max": r["spread"]}
        except Exception as exc:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                   "sample": f"failed: {exc!r}"}

    # ---------------- extra records (default N = 1 run only) -------------------------
    extras = None
    if (rank == 0 and world == 1 and not args.no_extras and args.config == "c3"
            and not (args.grid or args.nx or args.dtype or args.scaling)):
        del job
        torch.cuda.empty_cache()
        # insurance: should the run be cut short while the extras are being collected, the
        # headline numbers are at least in the stderr log
        print("[bench] headline (before extras): " + json.dumps({
            "value": value, "ms_per_step": ms_per_step, "steps": steps,
            "roofline_frac": roofline["frac"], "parity": parity,
            "e2e_value": (e2e or {}).get("value"),
            "cpu_baseline": (cpu or {}).get("value")}), file=sys.stderr, flush=True)
        try:
            extras = run_extras(args)
        except BaseException as exc:  # noqa: BLE001 -- never at the expense of the line
            extras = {"error": repr(exc)[:400]}

    if rank == 0:
        config = config_dict(cfg, w, world)
        line = {
            "metric": "MLUPS", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": config,
            "run": {
                "rank_shape": list(w["rank_shape"]),
                "halo_exchange": {"local": "periodic wrap written by the stage kernel",
                                  "peer": "stage kernels store their boundary planes into "
                                          "the neighbours' halos over NVLink (peer memory); "
                                          "one tiny kernel per stage all-gathers the energy "
                                          "sums over peer memory, is the inter-GPU barrier "
                                          "and advances the background",
                                  "nccl": "NCCL send/recv on a side stream"}[halo_mode],
                "launch": "CUDA graph (2 steps per replay)" if not args.no_graph
                          else "eager launches",
                "kernels_per_stage": kernels_per_stage,
                "full_step_mlups": value / NUM_STAGES,
            },
            "clocks": clk, "roofline": roofline, "parity": parity, "cpu_baseline": cpu,
            "e2e": e2e, "e2e_dropin": e2e_dropin, "extras": extras,
            "gpu_launches": launches_per_step * steps,
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


# ---------------------------------------------------------------------------
# extra records of the default run: the other single-GPU configurations of BASELINE.json and
# the diagnostics kernels, each in its own subprocess under a time limit, AFTER everything
# the headline line reports has been measured (a failure or timeout here costs one entry
# of "extras", never the line)
# ---------------------------------------------------------------------------

def _last_json_line(text):
    for ln in reversed(text.strip().splitlines()):
        ln = ln.strip()
        if ln.startswith("{") and ln.endswith("}"):
            try:
                return json.loads(ln)
            except ValueError:
                continue
    return None


def run_extras(args):
    import subprocess
    t_start = time.perf_counter()
    budget = float(args.extras_budget)
    out = {"what": "measured by the same bench.py in separate processes after the headline "
                   "configuration: `--config cX --steps S --warmup 3 --no-e2e --no-cpu` "
                   "(parity window kept) and `--diagnostics`; values are whole-GPU MLUPS",
           "budget_s": budget}
    jobs = [("c2", ["--config", "c2", "--steps", "40"]),
            ("c5", ["--config", "c5", "--steps", "10"]),
            ("c4", ["--config", "c4", "--steps", "6"]),
            ("diagnostics", ["--diagnostics"])]
    for name, extra in jobs:
        left = budget - (time.perf_counter() - t_start)
        if left < 20:
            out[name] = {"skipped": "extras budget exhausted"}
            continue
        cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1", "--warmup", "3",
               "--no-e2e", "--no-cpu", "--no-extras"] + extra
        t0 = time.perf_counter()
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=min(left, 120.0),
                                 cwd=ROOT)
            rec = _last_json_line(res.stdout)
            if rec is None:
                out[name] = {"error": f"rc={res.returncode}", "stderr_tail": res.stderr[-400:]}
                continue
        except subprocess.TimeoutExpired:
            out[name] = {"error": "timeout"}
            continue
        except Exception as exc:  # noqa: BLE001
            out[name] = {"error": repr(exc)}
            continue
        if name == "diagnostics":
            rec["seconds"] = time.perf_counter() - t0
            out[name] = rec
            continue
        roof = rec.get("roofline") or {}
        par = rec.get("parity") or {}
        summary = {
            "workload": (rec.get("config") or {}).get("workload"),
            "value": rec.get("value"), "unit": rec.get("unit"), "steps": rec.get("steps"),
            "ms_per_step": rec.get("ms_per_step"), "dtype": rec.get("dtype"),
            "roofline_frac": roof.get("frac"), "roofline_kernel": roof.get("kernel"),
            "frac_per_stage": roof.get("frac_per_stage"),
            "whole_stage": roof.get("whole_stage"),
            "parity_ok": par.get("ok"), "parity_max_rel": par.get("max_rel"),
            "parity_tol": par.get("tol"), "clocks": rec.get("clocks"),
            "gpu_launches": rec.get("gpu_launches"),
            "seconds": time.perf_counter() - t0,
        }
        out[name] = summary
    out["seconds"] = time.perf_counter() - t_start
    return out


def run_diagnostics(args, dev=None, timed=None, dft_backend=None):
    """(`dev`, `timed`, `dft_backend`: test hooks -- tests/test_emulated_api.py runs this
    function on host tensors against the CPU execution model.)  Times the diagnostics kernels of row f-4 at the headline lattice (512^3 fp64):
    FieldHistogrammer (bounds through psb_reduce + the generated two-histogram kernel) and
    PowerSpectra.bin_power (psb_bin_power), each against its algorithmic traffic (one read
    of the array) and against the tensor-operation formulation it replaces, which is also
    the correctness check.  CUDA events on the current stream, 3 warm-up + 10 timed calls."""
    import torch
    import pystella_b200 as ps
    from pystella_b200 import _kernels
    on_gpu = dev is None
    if on_gpu:
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
    n = args.grid or 512
    h, nb = 2, 432
    grid = (n, n, n)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0

    def cuda_timed(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    timed = timed or cuda_timed

    rec = {"lattice": list(grid), "dtype": "f64", "peak_gbs": peak}
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    gen = torch.Generator(device=dev).manual_seed(11)
    sites = float(n)**3

    def field_histogram():
        f = torch.randn((1,) + tuple(m + 2 * h for m in grid), dtype=torch.float64,
                        device=dev, generator=gen)
        fh = ps.FieldHistogrammer(decomp, nb, "float64", halo_shape=h)
        bounds = fh._bounds(f[0])
        ms_bounds = timed(lambda: fh._bounds(f[0]))
        gen_hist = lambda: ps.Histogrammer._device_histograms(  # noqa: E731
            fh, None, dict(f=f[0], **bounds))
        got = gen_hist()
        ms_hist = timed(gen_hist)
        eager = lambda: ps.Histogrammer._host_histograms(fh, dict(f=f[0], **bounds))  # noqa: E731
        ref = eager()
        ms_eager = timed(eager, reps=3, warm=1)
        diff = float((got - ref).abs().sum())
        return {
            "num_bins": nb, "histograms": 2,
            "bounds_ms": ms_bounds, "bounds_gbs": sites * 8 / ms_bounds / 1e6,
            "bounds_what": "one psb_reduce launch pair (max f, min f, max |f|, min |f|) + the "
                           "host read-back of 4 doubles",
            "kernel_ms": ms_hist, "kernel_gbs": sites * 8 / ms_hist / 1e6,
            "kernel_frac_of_peak": sites * 8 / ms_hist / 1e6 / peak,
            "tensor_ops_ms": ms_eager, "speedup_vs_tensor_ops": ms_eager / ms_hist,
            "sites_counted": [float(got[0].sum()), float(got[1].sum())],
            "abs_count_difference_vs_tensor_ops": diff,
            "ok": bool(got[0].sum() == sites and got[1].sum() == sites and diff <= 16),
            "kernel": "generated (NVRTC) two-histogram kernel, csrc/hist_accum.cuh",
        }

    def bin_power():
        fft = (ps.DFT(decomp, None, None, grid, "float64") if dft_backend is None
               else ps.DFT(decomp, None, None, grid, "float64", backend=dft_backend))
        L = 5.0
        dk = (2 * np.pi / L,) * 3
        spectra = ps.PowerSpectra(decomp, fft, dk, L**3)
        K = tuple(fft.shape(True))
        fk = torch.view_as_complex(torch.randn(K + (2,), dtype=torch.float64, device=dev,
                                               generator=gen))
        momenta = [fft.sub_k[m] for m in ("momenta_x", "momenta_y", "momenta_z")]
        run = lambda: _kernels.bin_power(fk, momenta, dk, spectra.bin_width, 3,  # noqa: E731
                                         n / 2, spectra.num_bins, True)
        got = run()
        ms_bin = timed(run)
        bins, mult, kmag = (t.to(dev) for t in spectra._host_tables())

        def eager_spec():
            k = fk.reshape(-1)
            wgt = mult * kmag**3 * (k.real**2 + k.imag**2)
            out = torch.zeros(spectra.num_bins, dtype=torch.float64, device=dev)
            out.index_add_(0, bins, wgt)
            return out
        ref = eager_spec()
        ms_eager = timed(eager_spec, reps=3, warm=1)
        modes = float(np.prod(K))
        err = float(((got - ref).abs().max() / ref.abs().max()))
        return {
            "kshape": list(K), "num_bins": spectra.num_bins,
            "kernel_ms": ms_bin, "kernel_gbs": modes * 16 / ms_bin / 1e6,
            "kernel_frac_of_peak": modes * 16 / ms_bin / 1e6 / peak,
            "tensor_ops_ms": ms_eager, "speedup_vs_tensor_ops": ms_eager / ms_bin,
            "max_rel_difference_vs_tensor_ops": err, "ok": bool(err < 1e-12),
            "kernel": "psb::bin_power_kernel<double, true> (csrc/diag.cu), algorithmic "
                      "traffic 16 B per mode",
        }

    def stencil_map():
        # the reference's own stencil benchmark (test/test_stencil.py:36-90 with
        # timing=True): bandwidth = (x.nbytes + y.nbytes) / time
        from pystella_b200.field import var
        xs, ys = var("x"), var("y")
        ii, jj, kk = var("i"), var("j"), var("k")
        md = {ys[ii, jj, kk]: (xs[ii + h + h, jj + h, kk + h] + xs[ii + h, jj + h + h, kk + h]
                               + xs[ii + h, jj + h, kk + h + h] + xs[ii - h + h, jj + h, kk + h]
                               + xs[ii + h, jj - h + h, kk + h] + xs[ii + h, jj + h, kk - h + h])}
        x = torch.rand(tuple(m + 2 * h for m in grid), dtype=torch.float64, device=dev,
                       generator=gen)
        y_t = torch.zeros(grid, dtype=torch.float64, device=dev)
        y_d = torch.zeros_like(y_t)
        tiled = ps.Stencil(md, h, prefetch_args=["x"], autotune=False)
        direct = ps.Stencil(md, h, prefetch_args=["x"], tile=False)
        default = ps.Stencil(md, h, prefetch_args=["x"])
        default(None, x=x, y=y_t)          # the first call picks the faster kernel
        ms_t = timed(lambda: tiled(None, x=x, y=y_t))
        ms_d = timed(lambda: direct(None, x=x, y=y_d))
        nbytes = x.numel() * 8 + y_t.numel() * 8
        return {
            "what": "six-point sum of test/test_stencil.py, h = 2: x staged in shared memory "
                    "vs read through L1/L2 (tile=False); the default object times both on "
                    "its first call and keeps the faster; bytes = x + y once",
            "tiled_ms": ms_t, "tiled_gbs": nbytes / ms_t / 1e6,
            "tiled_frac_of_peak": nbytes / ms_t / 1e6 / peak,
            "direct_ms": ms_d, "direct_gbs": nbytes / ms_d / 1e6,
            "default_picks": (default.tuning or {}).get("chosen"),
            "ok": bool(torch.equal(y_t, y_d)),
        }

    for name, section in (("field_histogram", field_histogram), ("bin_power", bin_power),
                          ("stencil_map", stencil_map)):
        try:
            rec[name] = section()
        except Exception as exc:  # noqa: BLE001
            rec[name] = {"error": repr(exc)[:400]}
        if on_gpu:
            torch.cuda.empty_cache()
    print(json.dumps(rec), flush=True)
    return rec


# ---------------------------------------------------------------------------
# parity of the multi-GPU path against the CPU oracle (outside the timed region)
# ---------------------------------------------------------------------------

def run_parity(job, cfg, w, world, rank, dist, nsteps=2, half=None, margin=22):
    """
    Restart from the seeded initial conditions, integrate `nsteps` steps on all ranks through
    the public API (fused stages, peer-memory halos, in-kernel background), gather the window
    of 2*half site planes centred on the boundary between the LAST and the FIRST rank ... no:
    between rank 0 and rank 1 (N > 1) or on the periodic boundary (N = 1), and compare it on
    rank 0 with oracle/oracle_c.c run on a sub-box (window + `margin` planes either side:
    information travels h = 2 planes per stage, 10 stages -> 20 planes; the sub-box is wrapped
    periodically in x, which only contaminates its outer 20 planes).  The oracle is fed the
    scale factor and Hubble rate the device used at every stage (read back, this is outside
    the timed region); for the single-field configurations on one box-sized lattice the
    background itself is checked against the full-lattice oracle in tests/.
    Semantics of the window as in pystella test/test_decomp.py:47-103 (halo data equal to the
    neighbour's interior, corners included).
    """
    import torch
    from oracle import c_oracle as CO
    import oracle as O
    ps = job.ps
    h, dx, dt = w["h"], w["dx"], w["dt"]
    ns = w["nscalars"]
    G = w["grid_shape"]
    if half is None:
        half = 16 if w["gw"] else 32
    half = min(half, w["rank_shape"][0])
    dev = job.dev

    # ---- device run, recording (a, hubble) at the start of every stage
    stepper = job.make_stepper()
    arrays = job.upload()
    expand = job.make_expansion()
    expand.attach(stepper)
    scal = []
    for _ in range(nsteps):
        for s in range(NUM_STAGES):
            st = expand.state.cpu().numpy()
            scal.append((float(st[0]), float(st[2])))
            stepper(s, a=expand.a, hubble=expand.hubble, **arrays)
            expand.step(s, stepper, dt)
    torch.cuda.synchronize()

    # ---- gather the window: global planes [xb - half, xb + half), xb = boundary
    xb = w["rank_shape"][0] if world > 1 else 0
    if world > 1:
        xb = workload(cfg, world, 1)["x_start"]
    names = ["f", "dfdt"] + (["hij", "dhijdt"] if w["gw"] else [])
    lo, hi = w["x_start"], w["x_start"] + w["rank_shape"][0]
    win = {}
    for name in names:
        t = arrays[name]
        out = torch.zeros((t.shape[0], 2 * half) + tuple(G[1:]), dtype=t.dtype, device=dev)
        for p in range(2 * half):
            gx = (xb - half + p) % G[0]
            if lo <= gx < hi:
                out[:, p] = t[:, gx - lo + h, h:-h, h:-h]
        if world > 1:
            dist.all_reduce(out)
        win[name] = out.cpu().numpy().astype(np.float64)
    del stepper, arrays, expand
    if rank != 0:
        return None

    # ---- oracle on the sub-box [xb - half - margin, xb + half + margin)
    nxs = 2 * (half + margin)
    x0 = xb - half - margin
    pad = (nxs + 2 * h, G[1] + 2 * h, G[2] + 2 * h)
    f = np.zeros((ns,) + pad)
    d = np.zeros((ns,) + pad)
    filled = np.zeros(nxs, dtype=bool)
    for r in range(world):
        wr = workload(cfg, world, r)
        rlo, rhi = wr["x_start"], wr["x_start"] + wr["rank_shape"][0]
        sel = [(p, (x0 + p) % G[0]) for p in range(nxs)]
        sel = [(p, gx) for p, gx in sel if rlo <= gx < rhi and not filled[p]]
        if not sel:
            continue
        fr, dr = initial_fields(wr)
        for p, gx in sel:
            f[:, p + h, h:-h, h:-h] = fr[:, gx - rlo + h, h:-h, h:-h]
            d[:, p + h, h:-h, h:-h] = dr[:, gx - rlo + h, h:-h, h:-h]
            filled[p] = True
        del fr, dr
    assert filled.all()
    N = (nxs, G[1], G[2])
    pot = CO.make_potential(ns, potential_terms(w))
    A, B, _ = O.LSRK_TABLES["LowStorageRK54"]
    kf, kd = np.zeros_like(f), np.zeros_like(f)
    lap = np.zeros((ns,) + N)
    if w["gw"]:
        hij, dhij = np.zeros((6,) + pad), np.zeros((6,) + pad)
        kh, kdh = np.zeros_like(hij), np.zeros_like(hij)
        dfdx, lap_hij = np.zeros((ns, 3) + N), np.zeros((6,) + N)
    it = iter(scal)
    for _ in range(nsteps):
        for s in range(NUM_STAGES):
            a, hub = next(it)
            for n in range(ns):
                CO.share_halos(f[n], h)
                CO.laplacian(f[n], lap[n], h, dx)
                if w["gw"]:
                    CO.gradient(f[n], dfdx[n], h, dx)
            if w["gw"]:
                for c in range(6):
                    CO.share_halos(hij[c], h)
                    CO.laplacian(hij[c], lap_hij[c], h, dx)
            CO.scalar_stage(f, d, kf, kd, lap, h, dt, A[s], B[s], a, hub, pot)
            if w["gw"]:
                CO.tensor_stage(hij, dhij, kh, kdh, lap_hij, dfdx, h, dt, A[s], B[s], hub)
    ref = {"f": f, "dfdt": d}
    if w["gw"]:
        ref.update(hij=hij, dhijdt=dhij)
    errs = {}
    for name in names:
        r = ref[name][:, h + margin:h + margin + 2 * half, h:-h, h:-h]
        g = win[name]
        scale = np.max(np.abs(r), axis=(1, 2, 3), keepdims=True)
        scale[scale == 0] = 1.0
        errs[name] = float(np.max(np.abs(g - r) / scale))
    tol = 1e-12 if cfg["dtype"] == "f64" else 2e-4
    max_rel = max(errs.values())
    return {"ok": bool(max_rel < tol), "max_rel": max_rel, "tol": tol, "per_array": errs,
            "steps": nsteps, "window_planes": 2 * half, "boundary_x": int(xb),
            "ranks_in_window": 2 if world > 1 else 1,
            "what": "CUDA (all ranks, public API, peer-memory halos) vs oracle/oracle_c.c on "
                    "rank 0: x-window straddling the rank-0/rank-1 boundary (the periodic "
                    "boundary at N = 1) after 2 full steps from the seeded initial "
                    "conditions; error = max |cuda - oracle| / max |oracle| per array"
                    + ("" if cfg["dtype"] == "f64" else
                       "; fp32 tolerance: the oracle is fp64 arithmetic on the same fp32 "
                       "initial data, the bound is the stencil's cancellation "
                       "eps32*|f|*7.5/dx^2*dt per stage")}


# ---------------------------------------------------------------------------
# end-to-end legs
# ---------------------------------------------------------------------------

def _max_over_ranks(sec, dist, world, dev):
    if world > 1:
        import torch
        t = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return sec


def run_e2e_resident(args, job, dist, world):
    """
    End to end through the public API as INTEGRATION.md recommends it: initial conditions
    uploaded from pinned host memory, the time loop of fused stages + DeviceExpansion called
    stage by stage from Python (no CUDA graph), after EVERY step the background and energies
    are read back to the host (what a run logs per step), final fields downloaded to pinned
    host memory.  Wall clock over the whole job, max over ranks; the three phases are also
    reported separately (the upload / download are paid once per job, not per step).
    """
    torch = job.torch
    w, dev = job.w, job.dev
    dt = w["dt"]
    steps = args.steps
    stepper = job.make_stepper()
    out_host = [torch.empty_like(p).pin_memory() for p in job.pin]

    def run(nsteps):
        t0 = time.perf_counter()
        arrays = job.upload()
        expand = job.make_expansion()
        expand.attach(stepper)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        log, d2h = [], 0
        for _ in range(nsteps):
            for s in range(NUM_STAGES):
                stepper(s, a=expand.a, hubble=expand.hubble, **arrays)
                expand.step(s, stepper, dt)
            rec = expand.get()          # device -> host: a, adot, hubble + energies
            log.append((rec["a"], rec["total"], rec["pressure"]))
            d2h += (5 + 2 * w["nscalars"] + 3) * 8
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out_host[0].copy_(arrays["f"], non_blocking=True)
        out_host[1].copy_(arrays["dfdt"], non_blocking=True)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        return log, d2h, (t1 - t0, t2 - t1, t3 - t2)

    run(1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    log, d2h_scalars, phases = run(steps)
    if world > 1:
        dist.barrier()
    sec = _max_over_ranks(time.perf_counter() - t0, dist, world, dev)
    gs = job.global_sites
    fb = job.field_bytes()
    return {"value": gs * NUM_STAGES * steps / sec / 1e6, "unit": UNIT, "steps": steps,
            "h2d_bytes_per_step": fb / steps,
            "d2h_bytes_per_step": fb / steps + d2h_scalars / steps,
            "h2d_ms": phases[0] * 1e3, "loop_ms_per_step": phases[1] * 1e3 / steps,
            "d2h_ms": phases[2] * 1e3,
            "value_loop_only": gs * NUM_STAGES * steps / phases[1] / 1e6,
            "what": "public API, device-resident loop (INTEGRATION.md): upload ICs from "
                    "pinned host memory, Python-driven fused stages + DeviceExpansion, "
                    "background and energies read back to the host after every step, final "
                    "fields downloaded; wall clock incl. all copies (upload/download once "
                    f"per job of {steps} steps; value_loop_only excludes them and does not "
                    "depend on --steps)",
            "a_final": float(log[-1][0])}


def run_e2e_dropin(args, job, dist, world):
    """
    The loop a pystella user writes (examples/scalar_preheating.py:258-266) through
    the public API: fields uploaded from pinned host memory, per stage the stepper,
    then FiniteDifferencer + Reduction (device -> host energies) feeding the HOST
    Expansion, whose scalars go back into the next stage; final fields downloaded.
    """
    torch, ps = job.torch, job.ps
    w, dev, decomp = job.w, job.dev, job.decomp
    h, dx, dt = w["h"], w["dx"], w["dt"]
    steps = max(1, args.steps)
    stepper = job.make_stepper(energy=False)
    derivs = ps.FiniteDifferencer(decomp, h, dx, rank_shape=w["rank_shape"])
    reduce_energy = ps.Reduction(decomp, job.scalar, halo_shape=h, callback=ps.get_rho_and_p,
                                 rank_shape=w["rank_shape"], grid_size=job.global_sites)
    lap_f = torch.zeros((w["nscalars"],) + tuple(w["rank_shape"]), dtype=job.tdtype,
                        device=dev)
    out_host = [torch.empty_like(p).pin_memory() for p in job.pin]

    def compute_energy(f, dfdt, a):
        derivs(None, fx=f, lap=lap_f)
        return reduce_energy(None, f=f, dfdt=dfdt, lap_f=lap_f, a=np.array(a))

    def run(nsteps):
        t0 = time.perf_counter()
        arrays = job.upload()
        f, dfdt = arrays["f"], arrays["dfdt"]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        energy = compute_energy(f, dfdt, 1.)
        expand = ps.Expansion(energy["total"], ps.LowStorageRK54, mpl=w["mpl"])
        d2h = 0
        for _ in range(nsteps):
            for s in range(NUM_STAGES):
                stepper(s, f=f, dfdt=dfdt, a=expand.a, hubble=expand.hubble)
                expand.step(s, energy["total"], energy["pressure"], dt)
                energy = compute_energy(f, dfdt, expand.a)
                d2h += 3 * 8
        t2 = time.perf_counter()    # the last energies are on the host: the loop has drained
        out_host[0].copy_(f, non_blocking=True)
        out_host[1].copy_(dfdt, non_blocking=True)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        return expand, d2h, (t1 - t0, t2 - t1, t3 - t2)

    run(1)  # warm-up (allocations, first-call paths)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    expand, d2h_scalars, phases = run(steps)
    if world > 1:
        dist.barrier()
    sec = _max_over_ranks(time.perf_counter() - t0, dist, world, dev)
    fb = job.field_bytes()
    return {"value": job.global_sites * NUM_STAGES * steps / sec / 1e6, "unit": UNIT,
            "steps": steps, "h2d_bytes_per_step": fb / steps,
            "d2h_bytes_per_step": fb / steps + d2h_scalars / steps,
            "h2d_ms": phases[0] * 1e3, "loop_ms_per_step": phases[1] * 1e3 / steps,
            "d2h_ms": phases[2] * 1e3,
            "value_loop_only": job.global_sites * NUM_STAGES * steps / phases[1] / 1e6,
            "what": "reference loop shape unchanged (examples/scalar_preheating.py:258-266): "
                    "upload ICs from pinned host memory, per stage fused stepper + "
                    "FiniteDifferencer + Reduction (energies read back to the host) + host "
                    "Expansion, download final fields; wall clock incl. all copies, "
                    f"upload/download ({fb / 1e9:.2f} GB each way) amortised over {steps} steps; "
                    "value_loop_only excludes the two bulk copies (the per-stage energy "
                    "read-backs stay in)",
            "a_final": float(expand.a[0])}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    elif a.diagnostics:
        run_diagnostics(a)
    else:
        run_b200_arm(a)
