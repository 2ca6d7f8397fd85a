"""
Parity at the sizes BASELINE.json's configurations are quoted on: the CUDA path through the
public API (fused stage kernels, in-kernel energy sums and background stage) against the C
oracle's reference loop (oracle/oracle_c.c: stepper -> expand.step -> share_halos + lap +
energies, examples/scalar_preheating.py:258-266) on the SAME seeded initial conditions as
bench.py -- config 2 (256^3 fp64, two full steps), one full step of config 3 (512^3 fp64),
and the bench's own window check (the path SCALE records run at N > 1) at N = 1 for every
configuration family.  fp64: <= 1e-12 of the array scale after every comparison.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402
import pystella_b200 as ps  # noqa: E402
import bench  # noqa: E402

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu]


def scale_err(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def _oracle_run(w, nsteps):
    h = w["h"]
    f0, d0 = bench.initial_fields(w)
    f, d = f0.astype(np.float64), d0.astype(np.float64)
    kf, kd = np.zeros_like(f), np.zeros_like(f)
    N = w["rank_shape"]
    lap = np.zeros((w["nscalars"],) + N)
    pot = CO.make_potential(w["nscalars"], bench.potential_terms(w))
    gs = float(np.prod(N))
    for n in range(w["nscalars"]):
        CO.share_halos(f[n], h)
        CO.laplacian(f[n], lap[n], h, w["dx"])
    en = CO.scalar_energy(f, d, lap, h, 1.0, gs, pot)
    # the background starts from the homogeneous-mode energy, as bench.py's DeviceExpansion
    adot = float(np.sqrt(8 * np.pi / 3 / w["mpl"]**2 * bench.background_energy(w)))
    state = np.array([1.0, adot, adot, 0.0, 0.0])
    A, B, _ = O.LSRK_TABLES["LowStorageRK54"]
    CO.preheating_steps(nsteps, A, B, f, d, kf, kd, lap, h, w["dx"], w["dt"], w["mpl"], gs,
                        pot, state, en, True)
    return f, d, state


@pytest.mark.parametrize("config,nsteps", [("c2", 2), ("c3", 1)])
def test_config_size_parity_fp64(config, nsteps):
    cfg = dict(bench.CONFIGS[config], name=config)
    w = bench.workload(cfg, 1, 0)
    h = w["h"]
    f_ref, d_ref, state_ref = _oracle_run(w, nsteps)

    decomp = ps.DomainDecomposition((1, 1, 1), h, rank_shape=w["rank_shape"])
    sector = ps.ScalarSector(1, potential=bench.potential_callable(w))
    stepper = ps.LowStorageRK54(sector, halo_shape=h, dt=w["dt"], dx=w["dx"], decomp=decomp,
                                MapKernel=ps.FusedStencilStage, energy=True)
    f0, d0 = bench.initial_fields(w)
    f, dfdt = torch.from_numpy(f0).cuda(), torch.from_numpy(d0).cuda()
    expand = ps.DeviceExpansion(bench.background_energy(w), ps.LowStorageRK54, mpl=w["mpl"],
                                grid_size=float(np.prod(w["grid_shape"])), nscalars=1)
    for _ in range(nsteps):
        for s in range(5):
            stepper(s, f=f, dfdt=dfdt, a=expand.a, hubble=expand.hubble)
            expand.step(s, stepper, w["dt"])
    torch.cuda.synchronize()
    got = expand.get()
    assert np.isclose(got["a"], state_ref[0], rtol=1e-13)
    assert np.isclose(got["adot"], state_ref[1], rtol=1e-12)
    assert np.isclose(got["hubble"], state_ref[2], rtol=1e-12)
    for g, r in ((f, f_ref), (dfdt, d_ref)):
        assert scale_err(O.interior(g[0].cpu().numpy(), h), O.interior(r[0], h)) < 1e-12


@pytest.mark.parametrize("config,grid", [("c2", None), ("c4", 128), ("c5", None)])
def test_bench_window_parity_single_gpu(config, grid):
    """bench.py's own parity block (x-window across the periodic boundary vs the oracle on a
    sub-box), run here so that a regression fails the suite and not only a bench line."""
    cfg = dict(bench.CONFIGS[config], name=config)
    if grid:
        cfg["grid"] = (grid,) * 3
    if config == "c5":
        cfg["grid"] = (128, 256, 256)
    w = bench.workload(cfg, 1, 0)
    decomp = ps.DomainDecomposition((1, 1, 1), w["h"], rank_shape=w["rank_shape"])
    job = bench.Job(ps, torch, cfg, w, decomp, torch.device("cuda", 0), {"xchunk": 0,
                    "variant": 0}, float(np.prod(w["grid_shape"])))
    par = bench.run_parity(job, cfg, w, 1, 0, None)
    assert par["ok"], par
    if cfg["dtype"] == "f64":
        assert par["max_rel"] < 1e-12


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 2e-4)])
@pytest.mark.parametrize("nscalars", [1, 2])
def test_persistent_variant_parity(dtype, tol, nscalars):
    """tune_variant 30 (csrc/fused_stage_pws.cuh: persistent CTAs, continuous pipeline, guided
    self-scheduling) through bench.py's parity block: window across the periodic boundary
    against the C oracle, energies and background included; and field-for-field against the
    default kernel (same per-site arithmetic: only the order of the energy sums differs)."""
    cfg = dict(bench.CONFIGS["c3"], name="c3", dtype=dtype, grid=(72, 40, 128))
    if nscalars == 2:
        cfg["system"] = "scalar2gw"     # exercises the two-field instance (+ tensor kernel)
        cfg["amp"] = 1e-2
    w = bench.workload(cfg, 1, 0)
    decomp = ps.DomainDecomposition((1, 1, 1), w["h"], rank_shape=w["rank_shape"])
    results = {}
    for variant in (30, 0):
        job = bench.Job(ps, torch, cfg, w, decomp, torch.device("cuda", 0),
                        {"xchunk": 0, "variant": variant}, float(np.prod(w["grid_shape"])))
        par = bench.run_parity(job, cfg, w, 1, 0, None, half=8, margin=22)
        assert par["ok"] and par["max_rel"] < tol, (variant, par)
        stepper = job.make_stepper()
        arrays = job.upload()
        expand = job.make_expansion()
        for s in range(5):
            stepper(s, a=expand.a, hubble=expand.hubble, **arrays)
            expand.step(s, stepper, w["dt"])
        torch.cuda.synchronize()
        results[variant] = ({k: v.cpu().numpy() for k, v in arrays.items()}, expand.get())
    for name in results[0][0]:
        a30, a0 = results[30][0][name], results[0][0][name]
        scale = np.max(np.abs(a0)) or 1.0
        assert np.max(np.abs(a30 - a0)) / scale < (1e-13 if dtype == "f64" else 1e-6), name
    assert np.isclose(results[30][1]["a"], results[0][1]["a"], rtol=1e-13)


@pytest.mark.parametrize("system,dtype", [("scalar1", "f64"), ("scalar1", "f32"),
                                          ("scalar2gw", "f32"), ("scalar2gw", "f64")])
def test_programmatic_dependent_launch_is_bit_identical(system, dtype):
    """``tune={"pdl": True}`` (PSB_LAUNCH_PDL, include/pystella_b200.h): the stage kernels are
    scheduled while the previous kernel drains and block in ``griddepcontrol.wait`` before they
    touch memory.  Same kernels, same order of every sum -> the fields and the background
    must be BIT-identical to plain stream order, for launches issued back to back (no host
    synchronisation in the loop) and for the CUDA-graph replay bench.py times; the plain
    result is the one the oracle tests above pin."""
    cfg = dict(bench.CONFIGS["c3"], name="c3", dtype=dtype, grid=(96, 64, 128))
    if system == "scalar2gw":
        cfg["system"] = "scalar2gw"     # generated two-field stage + tensor kernel
        cfg["amp"] = 1e-2
    w = bench.workload(cfg, 1, 0)
    decomp = ps.DomainDecomposition((1, 1, 1), w["h"], rank_shape=w["rank_shape"])
    dev = torch.device("cuda", 0)

    def run(pdl, graph):
        job = bench.Job(ps, torch, cfg, w, decomp, dev, {"xchunk": 0, "variant": 0,
                        "pdl": pdl}, float(np.prod(w["grid_shape"])))
        stepper = job.make_stepper()
        assert stepper.launch_flags == (1 if pdl else 0)
        arrays = job.upload()
        expand = job.make_expansion()
        expand.attach(stepper)

        def one_step():
            for s in range(5):
                stepper(s, a=expand.a, hubble=expand.hubble, **arrays)
                expand.step(s, stepper, w["dt"])
        if graph:
            g = ps.capture_steps(one_step, steps_per_replay=2, warmup=1, stepper=stepper)
            g.replay()
            g.replay()
        else:
            for _ in range(6):
                one_step()
        torch.cuda.synchronize()
        return ({k: v.cpu().numpy().copy() for k, v in arrays.items()},
                expand.state.cpu().numpy().copy())

    for graph in (False, True):
        ref_arrays, ref_state = run(False, graph)
        got_arrays, got_state = run(True, graph)
        for name in ref_arrays:
            assert np.isfinite(ref_arrays[name]).all(), name
            assert np.array_equal(got_arrays[name], ref_arrays[name]), (name, graph)
        assert np.array_equal(got_state, ref_state), graph
