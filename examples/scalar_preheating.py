"""
Preheating of scalar fields in an expanding FLRW background: the time loop of the
reference's examples/scalar_preheating.py (sectors, stepper, derivatives, energy
reduction, expansion) on pystella_b200, with the reference's initial conditions: WKB
vacuum fluctuations drawn by RayleighGenerator on top of the homogeneous background
(the distributed DFT covers slab and pencil decompositions).
Projectors and HDF5 output are not provided.

    python examples/scalar_preheating.py -grid 64 64 64 --end-time 1            # 1 GPU
    torchrun --nproc-per-node 4 examples/scalar_preheating.py -grid 256 256 256 -proc 4 1 1
Options: --reference-loop (derivs + Reduction + host Expansion per stage, exactly the
reference's loop) instead of the default device-resident loop; -gws adds the tensor sector.
"""
import argparse
import os
import sys
from time import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pystella_b200 as ps  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--grid-shape", "-grid", type=int, nargs=3, default=(128, 128, 128))
parser.add_argument("--proc-shape", "-proc", type=int, nargs=3, default=(1, 1, 1))
parser.add_argument("--dtype", default="float64", choices=["float64", "float32"])
parser.add_argument("--halo-shape", type=int, default=2, metavar="h")
parser.add_argument("--box-dim", "-box", type=float, nargs=3, default=(5, 5, 5))
parser.add_argument("--kappa", type=float, default=1/10)
parser.add_argument("--mpl", type=float, default=1)
parser.add_argument("--mphi", type=float, default=1.20e-6)
parser.add_argument("--mchi", type=float, default=0.)
parser.add_argument("--gsq", type=float, default=2.5e-7)
parser.add_argument("--sigma", type=float, default=0.)
parser.add_argument("--lambda4", type=float, default=0.)
parser.add_argument("--end-time", "-end-t", type=float, default=20)
parser.add_argument("--end-scale-factor", "-end-a", type=float, default=20)
parser.add_argument("--gravitational-waves", "-gws", action="store_true")
parser.add_argument("--reference-loop", action="store_true")
parser.add_argument("--initial-conditions", "-ic", choices=["rayleigh", "white"],
                    default=None, help="default: rayleigh (WKB vacuum modes, like the reference)")
parser.add_argument("--seed", type=int, default=49279)
parser.add_argument("--spectra", action="store_true",
                    help="print the field power spectra at the end")
parser.add_argument("--output", metavar="NAME", default=None,
                    help="write energies, field statistics and spectra (every 5%% growth "
                         "of the scale factor) to NAME.npz, like the reference's OutputFile")


def init_rayleigh(p, decomp, f, dfdt, f0, potential, energy, expand, backend=None):
    """WKB vacuum fluctuations into (the interiors of) ``f`` and ``dfdt``: Hubble correction
    to the effective masses from the background, then one draw per field (reference
    example, lines 213-236).  ``backend="fftw"`` works on host arrays."""
    from pystella_b200.field import evaluate, var
    nscalars = len(f0)
    addot = expand.addot_friedmann_2(expand.a, energy["total"], energy["pressure"])
    hubble_correction = - addot / expand.a
    fields = [var("f0")[i] for i in range(nscalars)]
    d2Vd2f = [ps.diff(potential(fields), field, field) for field in fields]
    eff_mass = [float(evaluate(x, {"f0": f0}) + hubble_correction[0]) for x in d2Vd2f]
    fft = ps.DFT(decomp, None, None, p.grid_shape, p.dtype, backend=backend)
    dk = tuple(2 * np.pi / Li for Li in p.box_dim)
    volume = float(np.prod(p.box_dim))
    modes = ps.RayleighGenerator(None, fft, dk, volume, seed=p.seed * (decomp.rank + 1))
    for fld in range(nscalars):
        modes.init_WKB_fields(f[fld], dfdt[fld], norm=p.mphi**2,
                              omega_k=lambda k, fld=fld: np.sqrt(k**2 + eff_mass[fld]),
                              hubble=expand.hubble[0])
    return eff_mass


def make_rho_diagnostics(decomp, scalar_sector, mpl, halo_shape, rank_shape, dtype):
    """The reference example's energy-density diagnostics (lines 139-146): a
    FieldHistogrammer with 1000 bins and the element-wise map
    ``rho = T_00 / (3 mpl^2 hubble^2 / 8 pi)`` (conformal-time quantities)."""
    hist = ps.FieldHistogrammer(decomp, 1000, dtype, rank_shape=rank_shape)
    a_sq_rho = 3 * mpl**2 * ps.Field("hubble", indices=[])**2 / 8 / np.pi
    rho_dict = {ps.Field("rho"): scalar_sector.stress_tensor(0, 0) / a_sq_rho}
    compute_rho = ps.ElementWiseMap(rho_dict, halo_shape=halo_shape, rank_shape=rank_shape,
                                    dtype=dtype)
    return hist, compute_rho


def make_potential(p):
    def potential(f):
        phi, chi = f[0], f[1]
        unscaled = (p.mphi**2 / 2 * phi**2 + p.mchi**2 / 2 * chi**2
                    + p.gsq / 2 * phi**2 * chi**2 + p.sigma / 2 * phi * chi**2
                    + p.lambda4 / 4 * chi**4)
        return unscaled / p.mphi**2
    return potential


def main(argv=None):
    """-> dict with the final time, scale factor, Friedmann constraint and energies."""
    p = parser.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = ps.choose_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    p.grid_shape = tuple(p.grid_shape)
    p.grid_size = float(np.prod(p.grid_shape))
    p.proc_shape = tuple(p.proc_shape)
    h = p.halo_shape
    dtype = getattr(torch, p.dtype)
    p.dx = tuple(Li / Ni for Li, Ni in zip(p.box_dim, p.grid_shape))
    dt = p.kappa * min(p.dx)

    nscalars = 2
    f0 = [.193 * p.mpl, 0]
    df0 = [-.142231 * p.mpl, 0]
    Stepper = ps.LowStorageRK54

    decomp = ps.DomainDecomposition(p.proc_shape, h, grid_shape=p.grid_shape)
    rank_shape = decomp.rank_shape
    pencil_shape = tuple(ni + 2 * h for ni in rank_shape)
    derivs = ps.FiniteDifferencer(decomp, h, p.dx, rank_shape=rank_shape)

    potential = make_potential(p)
    scalar_sector = ps.ScalarSector(nscalars, potential=potential)
    sectors = [scalar_sector]
    if p.gravitational_waves:
        sectors += [ps.TensorPerturbationSector([scalar_sector])]

    device_loop = not (p.reference_loop or p.gravitational_waves)
    slab = p.proc_shape[1] == 1
    fused = dict(MapKernel=ps.FusedStencilStage, dx=p.dx, decomp=decomp) if slab else {}
    stepper = Stepper(sectors, halo_shape=h, rank_shape=rank_shape, dt=dt,
                      energy=device_loop, **fused)

    reduce_energy = ps.Reduction(decomp, scalar_sector, halo_shape=h,
                                 callback=ps.get_rho_and_p, rank_shape=rank_shape,
                                 grid_size=p.grid_size)
    statistics = ps.FieldStatistics(decomp, h, rank_shape=rank_shape,
                                    grid_size=p.grid_size)

    ic = p.initial_conditions or "rayleigh"
    f = torch.empty((nscalars,) + pencil_shape, dtype=dtype, device=device)
    dfdt = torch.empty_like(f)
    for i in range(nscalars):           # field means
        f[i] = f0[i]
        dfdt[i] = df0[i]
    dfdx = torch.zeros((nscalars, 3) + rank_shape, dtype=dtype, device=device)
    lap_f = torch.zeros((nscalars,) + rank_shape, dtype=dtype, device=device)
    if p.gravitational_waves:
        hij = torch.zeros((6,) + pencil_shape, dtype=dtype, device=device)
        dhijdt = torch.zeros_like(hij)
        lap_hij = torch.zeros((6,) + rank_shape, dtype=dtype, device=device)
    else:
        hij = dhijdt = lap_hij = None

    def compute_energy(a):
        if p.gravitational_waves:
            derivs(None, fx=f, lap=lap_f, grd=dfdx)
        else:
            derivs(None, fx=f, lap=lap_f)
        return reduce_energy(None, f=f, dfdt=dfdt, lap_f=lap_f, a=np.array(a))

    # energy of the background fields initialises the expansion
    energy = compute_energy(1.)
    expand = ps.Expansion(energy["total"], Stepper, mpl=p.mpl)
    if ic == "rayleigh":
        init_rayleigh(p, decomp, f, dfdt, f0, potential, energy, expand)
        for i in range(nscalars):
            f[i] += f0[i]
            dfdt[i] += df0[i]
    else:
        gen = torch.Generator(device=device).manual_seed(p.seed * (decomp.rank + 1))
        for i in range(nscalars):
            f[i] += 1e-6 * torch.randn(pencil_shape, dtype=dtype, device=device,
                                       generator=gen)
            dfdt[i] += 1e-6 * torch.randn(pencil_shape, dtype=dtype, device=device,
                                          generator=gen)

    # re-initialise energy and expansion
    energy = compute_energy(expand.a[0])
    if device_loop:
        expand = ps.DeviceExpansion(energy["total"], Stepper, mpl=p.mpl,
                                    grid_size=p.grid_size, decomp=decomp, nscalars=nscalars)
    else:
        expand = ps.Expansion(energy["total"], Stepper, mpl=p.mpl)
        if p.gravitational_waves:
            derivs(None, fx=hij, lap=lap_hij)

    # output as in the reference example (lines 131-190)
    out = spectra = projector = hist = compute_rho = None
    if p.output is not None:
        if decomp.rank == 0:
            out = ps.OutputFile(name=p.output, runfile=__file__, grid_shape=p.grid_shape,
                                proc_shape=p.proc_shape, dtype=p.dtype, mphi=p.mphi,
                                gsq=p.gsq, initial_conditions=ic)
        if True:
            dk = tuple(2 * np.pi / Li for Li in p.box_dim)
            fft_out = ps.DFT(decomp, None, None, p.grid_shape, p.dtype)
            spectra = ps.PowerSpectra(decomp, fft_out, dk, float(np.prod(p.box_dim)))
            projector = ps.Projector(fft_out, h, dk, p.dx)
            hist, compute_rho = make_rho_diagnostics(decomp, scalar_sector, p.mpl, h,
                                                     rank_shape, p.dtype)
    a_last_spec = [.1]

    def output(t, a, hubble, energy, constraint):
        f_stats = statistics(f)
        if out is not None:
            out.output("energy", t=t, a=a, hubble=hubble / a, constraint=constraint,
                       eos=energy["pressure"] / energy["total"],
                       **{k: np.asarray(v) for k, v in energy.items()})
            out.output("statistics/f", t=t, a=a, **f_stats)
        if spectra is not None and a / a_last_spec[0] >= 1.05:
            a_last_spec[0] = a
            # energy density in units of the critical density, its histogram and spectrum
            # (reference lines 155-163, 171)
            if not p.gravitational_waves:
                derivs(None, fx=f, grd=dfdx)
            rho = torch.empty(rank_shape, dtype=dtype, device=device)
            compute_rho(None, a=np.array(a), hubble=np.array(hubble), rho=rho, f=f,
                        dfdt=dfdt, dfdx=dfdx)
            rho_hist = hist(rho)
            spec_out = {"scalar": spectra(f), "rho": spectra(rho)}
            if p.gravitational_waves:
                spec_out["gw"] = spectra.gw(dhijdt, projector, hubble)
            if out is not None:
                out.output("rho_histogram", t=t, a=a, **rho_hist)
                out.output("spectra", t=t, a=a, **spec_out)

    t, step_count = 0., 0
    if decomp.rank == 0:
        print("time\t\tscale factor\tconstraint\tms/step")
    start = time()
    a_now = 1.
    while t < p.end_time and a_now < p.end_scale_factor:
        for s in range(stepper.num_stages):
            if device_loop:
                stepper(s, f=f, dfdt=dfdt, a=expand.a, hubble=expand.hubble)
                expand.step(s, stepper, dt)
            else:
                stepper(s, queue=None, a=expand.a, hubble=expand.hubble, f=f, dfdt=dfdt,
                        dfdx=dfdx, lap_f=lap_f, hij=hij, dhijdt=dhijdt, lap_hij=lap_hij,
                        filter_args=True)
                expand.step(s, energy["total"], energy["pressure"], dt)
                energy = compute_energy(expand.a)
                if p.gravitational_waves:
                    derivs(None, fx=hij, lap=lap_hij)
        t += dt
        step_count += 1
        if step_count % 4 == 0 or t >= p.end_time:
            if device_loop:
                rec = expand.get()
                a_now, hubble_now = float(rec["a"]), float(rec["hubble"])
                # energy of the current state for the output (stand-alone kernels)
                energy = compute_energy(a_now)
                constraint = float(expand.constraint(energy["total"]))
            else:
                a_now, hubble_now = float(expand.a[0]), float(expand.hubble[0])
                constraint = float(expand.constraint(energy["total"]))
            if p.output is not None:
                output(t, a_now, hubble_now, energy, constraint)
            if decomp.rank == 0:
                ms = (time() - start) * 1e3 / step_count
                print(f"{t:<15.3f}\t{a_now:<15.6f}\t{constraint:<.3e}\t{ms:<.3f}")

    stats = statistics(f)
    if decomp.rank == 0:
        print("field means", stats["mean"], "variances", stats["variance"])
    if p.spectra:
        fft = ps.DFT(decomp, None, None, p.grid_shape, p.dtype)
        dk = tuple(2 * np.pi / Li for Li in p.box_dim)
        spectra = ps.PowerSpectra(decomp, fft, dk, float(np.prod(p.box_dim)))
        scalar_spectra = spectra(f)
        if decomp.rank == 0:
            print("k bins", spectra.bin_width * np.arange(spectra.num_bins))
            print("scalar spectra", scalar_spectra)
    if out is not None:
        out.close()
    if decomp.rank == 0:
        print("Simulation complete")
    result = {"t": t, "steps": step_count, "a": a_now, "constraint": constraint,
              "initial_conditions": ic}
    if world > 1:
        dist.barrier()
        os._exit(0)
    return result


if __name__ == "__main__":
    main()
