"""Test-side helpers (may import the oracle; product code may not)."""
import numpy as np
import torch

import oracle as O


class HostOps:
    """
    Oracle-backed halo operations for *host* tensors, injected into
    DomainDecomposition.host_ops by the gloo tests so that the rank arithmetic and
    message ordering of share_halos can be exercised without a GPU.
    """

    def wrap(self, fx, halo_shape, mask):
        h = [hh if (mask >> ax) & 1 else 0 for ax, hh in enumerate(halo_shape)]
        arr = fx.numpy()
        N = [n - 2 * hh for n, hh in zip(arr.shape, halo_shape)]
        hx, hy, hz = halo_shape
        if h[2]:
            arr[:, :, :hz] = arr[:, :, N[2]:N[2] + hz]
            arr[:, :, N[2] + hz:] = arr[:, :, hz:2 * hz]
        if h[0]:
            arr[:hx] = arr[N[0]:N[0] + hx]
            arr[N[0] + hx:] = arr[hx:2 * hx]
        if h[1]:
            arr[:, :hy] = arr[:, N[1]:N[1] + hy]
            arr[:, N[1] + hy:] = arr[:, hy:2 * hy]

    def pack(self, fx, buf, halo_shape, axis, unpack):
        assert axis == 1
        hy = halo_shape[1]
        Ny = fx.shape[1] - 2 * hy
        if not unpack:
            buf[1].copy_(fx[:, Ny:Ny + hy, :])
            buf[0].copy_(fx[:, hy:2 * hy, :])
        else:
            fx[:, :hy, :] = buf[1]
            fx[:, Ny + hy:, :] = buf[0]


def periodic_global(rng, grid_shape, h, dtype=np.float64):
    """Globally periodic padded array as in test/test_decomp.py:61-76."""
    h = (h,) * 3 if isinstance(h, int) else tuple(h)
    data = rng.random(tuple(N + 2 * hi for N, hi in zip(grid_shape, h))).astype(dtype)
    if h[0] > 0:
        data[:h[0], :, :] = data[-2 * h[0]:-h[0], :, :]
        data[-h[0]:, :, :] = data[h[0]:2 * h[0], :, :]
    if h[1] > 0:
        data[:, :h[1], :] = data[:, -2 * h[1]:-h[1], :]
        data[:, -h[1]:, :] = data[:, h[1]:2 * h[1], :]
    if h[2] > 0:
        data[:, :, :h[2]] = data[:, :, -2 * h[2]:-h[2]]
        data[:, :, -h[2]:] = data[:, :, h[2]:2 * h[2]]
    return data


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def scale_err(got, ref):
    """max |got-ref| / max |ref|  (array-scale relative error)."""
    ref = np.asarray(ref)
    return float(np.max(np.abs(np.asarray(got) - ref)) / max(np.max(np.abs(ref)), 1e-300))


__all__ = ["HostOps", "periodic_global", "to_dev", "scale_err", "O"]


# ---------------------------------------------------------------------------------------
# the reference's preheating example, end to end on the oracle
# ---------------------------------------------------------------------------------------

def example_parameters(golden):
    """Defaults of examples/scalar_preheating.py as extracted by make_golden.py."""
    d = golden["example"]["defaults"]
    num = lambda s: float(eval(s, {"__builtins__": {}}))  # noqa: E731  ("1/10", "1.20e-6")
    mpl = num(d["mpl"])
    env = {"__builtins__": {}, "p": type("P", (), {"mpl": mpl})}
    return {
        "h": int(d["halo-shape"]), "box": eval(d["box-dim"], {"__builtins__": {}}),
        "kappa": num(d["kappa"]), "mpl": mpl, "mphi": num(d["mphi"]), "mchi": num(d["mchi"]),
        "gsq": num(d["gsq"]), "sigma": num(d["sigma"]), "lambda4": num(d["lambda4"]),
        "f0": list(eval("[" + golden["example"]["f0"] + "]", env)),
        "df0": list(eval("[" + golden["example"]["df0"] + "]", env)),
    }


def example_potential_terms(p):
    """examples/scalar_preheating.py:98-105 as polynomial terms (coef, (pow_phi, pow_chi))."""
    m2 = p["mphi"]**2
    terms = [(p["mphi"]**2 / 2 / m2, (2, 0)), (p["mchi"]**2 / 2 / m2, (0, 2)),
             (p["gsq"] / 2 / m2, (2, 2)), (p["sigma"] / 2 / m2, (1, 2)),
             (p["lambda4"] / 4 / m2, (0, 4))]
    return [(c, pw) for c, pw in terms if c != 0]


def example_initial_conditions(p, grid_shape, seed):
    """
    examples/scalar_preheating.py:193-236 on the oracle: homogeneous background ->
    Hubble correction to the effective masses -> WKB Rayleigh modes -> re-initialised
    expansion.  The uniform deviates come from NumPy (the reference's Threefry stream is
    not available), so results agree with the reference statistically only.
    -> dict(f, dfdt (padded, halos NOT yet shared), a, adot, hubble, energy, dx, dt, ...)
    """
    from oracle import c_oracle as OC, fourier_oracle as F
    ns, h = 2, p["h"]
    N = tuple(grid_shape)
    dx = tuple(Li / Ni for Li, Ni in zip(p["box"], N))
    dk = tuple(2 * np.pi / Li for Li in p["box"])
    volume = float(np.prod(p["box"]))
    dt = p["kappa"] * min(dx)
    V = example_potential_terms(p)
    pot = OC.make_potential(ns, V)
    pad = tuple(n + 2 * h for n in N)
    f, d = np.zeros((ns,) + pad), np.zeros((ns,) + pad)
    lap = np.zeros((ns,) + N)
    grid_size = float(np.prod(N))
    for i in range(ns):
        f[i], d[i] = p["f0"][i], p["df0"][i]

    def energy(a):
        for i in range(ns):
            OC.share_halos(f[i], h)
            OC.laplacian(f[i], lap[i], h, dx)
        return OC.scalar_energy(f, d, lap, h, a, grid_size, pot)

    def rho_p(en):
        kin, potl, grad = en[:ns].sum(), en[ns], en[ns + 1:].sum()
        return kin + potl + grad, kin - grad / 3 - potl        # sectors.py:211-229

    total, pressure = rho_p(energy(1.))
    ex = O.Expansion(total, "LowStorageRK54", mpl=p["mpl"])
    addot = ex.addot_friedmann_2(ex.a, total, pressure)
    correction = float((-addot / ex.a)[0])
    m2 = p["mphi"]**2
    f0 = p["f0"]
    d2V = [(p["mphi"]**2 + p["gsq"] * f0[1]**2) / m2,              # d2V/dphi2 at f0
           (p["mchi"]**2 + p["gsq"] * f0[0]**2 + p["sigma"] * f0[0]
            + 3 * p["lambda4"] * f0[1]**2) / m2]                   # d2V/dchi2 at f0
    eff_mass = [m + correction for m in d2V]

    rng = np.random.default_rng(seed)
    kshape = (N[0], N[1], N[2] // 2 + 1)
    for fld in range(ns):
        rands = rng.random((4,) + kshape)
        rands[rands == 0] = 0.5
        fx, dfx = F.init_wkb_fields(
            rands, N, dk, volume, norm=p["mphi"]**2,
            omega_k=lambda k, fld=fld: np.sqrt(k**2 + eff_mass[fld]),
            hubble=float(ex.hubble[0]))
        O.interior(f[fld], h)[...] = fx
        O.interior(d[fld], h)[...] = dfx
    for i in range(ns):
        f[i] += p["f0"][i]
        d[i] += p["df0"][i]
    en = energy(float(ex.a[0]))
    total, pressure = rho_p(en)
    ex = O.Expansion(total, "LowStorageRK54", mpl=p["mpl"])
    return {"f": f, "dfdt": d, "lap": lap, "energies": en, "expansion": ex, "pot": pot,
            "V_terms": V, "dx": dx, "dt": dt, "grid_size": grid_size, "rho_p": rho_p,
            "eff_mass": eff_mass}


def example_run_oracle(p, ic, end_time):
    """The time loop of examples/scalar_preheating.py:258-272 on the C oracle;
    -> (number of steps, a, constraint (expansion.py:159-176), total energy)."""
    from oracle import c_oracle as OC
    f, d, lap, en, ex = ic["f"], ic["dfdt"], ic["lap"], ic["energies"], ic["expansion"]
    A, B, _ = O.LSRK_TABLES["LowStorageRK54"]
    state = np.array([ex.a[0], ex.adot[0], ex.hubble[0], 0., 0.])
    kf, kd = np.zeros_like(f), np.zeros_like(d)
    nsteps, t = 0, 0.
    while t < end_time:
        t += ic["dt"]
        nsteps += 1
    OC.preheating_steps(nsteps, A, B, f, d, kf, kd, lap, p["h"], ic["dx"], ic["dt"],
                        p["mpl"], ic["grid_size"], ic["pot"], state, en)
    total, _ = ic["rho_p"](en)
    a, adot = state[0], state[1]
    constraint = abs(np.sqrt(8 * np.pi * a**2 / 3 / p["mpl"]**2 * total) * a / adot - 1)
    return nsteps, a, constraint, total


def gw_stage_case(N, h, dtype, nstages, rng, to_device):
    """Two coupled scalars + tensor sector under FusedStencilStage (config 4's stage) on the
    lattice N: `nstages` LSRK54 stages through the public API against the oracle's scalar_rhs
    + gw_rhs + lsrk_stage.  `to_device` uploads a NumPy array (a copy)."""
    import pystella_b200 as ps
    from test_gpu_api import _example_potential
    tol = 1e-12 if dtype == np.float64 else 3e-5
    pad = tuple(n + 2 * h for n in N)
    dx = tuple(5 / n for n in N)
    dt = 0.1 * min(dx)
    potential, terms = _example_potential(mphi=1.0, gsq=0.3)
    f0 = (0.2 * rng.standard_normal((2,) + pad)).astype(dtype)
    d0 = (0.2 * rng.standard_normal((2,) + pad)).astype(dtype)
    h0 = (1e-2 * rng.standard_normal((6,) + pad)).astype(dtype)
    dh0 = (1e-2 * rng.standard_normal((6,) + pad)).astype(dtype)
    a, hubble = 1.1, 0.3
    F, D, Hh, DH = (x.astype(np.float64) for x in (f0, d0, h0, dh0))
    kF, kD, kH, kDH = (np.zeros_like(x) for x in (F, D, Hh, DH))
    A_, B_, _ = O.LSRK_TABLES["LowStorageRK54"]
    for s_ in range(nstages):
        lapf, grd, laph = [], [], []
        for n in range(2):
            O.share_halos(F[n], h)
            lapf.append(O.laplacian(F[n], h, dx))
            grd.append(O.gradient(F[n], h, dx))
        for c in range(6):
            O.share_halos(Hh[c], h)
            laph.append(O.laplacian(Hh[c], h, dx))
        Fi = [O.interior(F[n], h) for n in range(2)]
        Di = [O.interior(D[n], h) for n in range(2)]
        Hi = [O.interior(Hh[c], h) for c in range(6)]
        DHi = [O.interior(DH[c], h) for c in range(6)]
        rhs = [np.array(r) for r in O.scalar_rhs(Fi, Di, lapf, a, hubble, terms)]
        rhs += [np.array(r) for r in O.gw_rhs(DHi, laph, np.stack(grd), hubble)]
        unk, tmp = [], []
        for n in range(2):
            unk += [Fi[n], Di[n]]
            tmp += [O.interior(kF[n], h), O.interior(kD[n], h)]
        for c in range(6):
            unk += [Hi[c], DHi[c]]
            tmp += [O.interior(kH[c], h), O.interior(kDH[c], h)]
        O.lsrk_stage(A_[s_], B_[s_], dt, unk, tmp, rhs)

    decomp = ps.DomainDecomposition((1, 1, 1), h, N)
    scalar_sector = ps.ScalarSector(2, potential=potential)
    gw_sector = ps.TensorPerturbationSector([scalar_sector])
    stepper = ps.LowStorageRK54([scalar_sector, gw_sector], halo_shape=h, rank_shape=N, dt=dt,
                                MapKernel=ps.FusedStencilStage, dx=dx, decomp=decomp)
    derivs = ps.FiniteDifferencer(decomp, h, dx, rank_shape=N)
    f, dfdt, hij, dhijdt = (to_device(x) for x in (f0, d0, h0, dh0))
    lap_f = torch.zeros((2,) + N, dtype=f.dtype, device=f.device)
    dfdx = torch.zeros((2, 3) + N, dtype=f.dtype, device=f.device)
    lap_hij = torch.zeros((6,) + N, dtype=f.dtype, device=f.device)
    for s_ in range(nstages):
        derivs(None, fx=f, lap=lap_f, grd=dfdx)
        derivs(None, fx=hij, lap=lap_hij)
        stepper(s_, queue=None, a=np.array(a), hubble=np.array(hubble), f=f, dfdt=dfdt,
                dfdx=dfdx, lap_f=lap_f, hij=hij, dhijdt=dhijdt, lap_hij=lap_hij,
                filter_args=True)
    for got, ref in ((f, F), (dfdt, D), (hij, Hh), (dhijdt, DH)):
        for c in range(ref.shape[0]):
            assert scale_err(O.interior(got[c].cpu().numpy(), h), O.interior(ref[c], h)) < tol, \
                (N, h, dtype, nstages, c)
