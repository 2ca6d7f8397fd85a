"""
TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference hot path.

Every function cites the reference lines (relative to /root/reference) whose
arithmetic it restates.  Arrays follow the reference layout: padded fields have
shape ``(Nx+2h, Ny+2h, Nz+2h)`` in C order (z fastest), derivative outputs have
shape ``(Nx, Ny, Nz)``.
"""

import numpy as np

__all__ = [
    "GRAD_COEFS", "LAP_COEFS", "first_eigenvalue", "second_eigenvalue",
    "interior", "share_halos", "laplacian", "gradient", "partial", "divergence",
    "LSRK_TABLES", "RK_TABLES", "lsrk_stage", "rk_stage", "rk_rhs_copy",
    "scalar_rhs", "gw_rhs", "tensor_index",
    "scalar_energy", "get_rho_and_p", "field_statistics",
    "Expansion", "rank_shape_start", "rank_id", "rank_tuple_of",
    "get_errs", "reference_schedule_stage", "poly_eval", "poly_diff",
]


# --------------------------------------------------------------------------
# finite-difference tables -- pystella/derivs.py:127-131 and :160-165
# --------------------------------------------------------------------------

GRAD_COEFS = {
    1: {1: 1/2},
    2: {1: 8/12, 2: -1/12},
    3: {1: 45/60, 2: -9/60, 3: 1/60},
    4: {1: 672/840, 2: -168/840, 3: 32/840, 4: -3/840},
}

LAP_COEFS = {
    1: {0: -2, 1: 1},
    2: {0: -30/12, 1: 16/12, 2: -1/12},
    3: {0: -490/180, 1: 270/180, 2: -27/180, 3: 2/180},
    4: {0: -14350/5040, 1: 8064/5040, 2: -1008/5040, 3: 128/5040, 4: -9/5040},
}


def first_eigenvalue(h, k, dx):
    """pystella/derivs.py:141-157 (FirstCenteredDifference.get_eigenvalues)."""
    th = k * dx
    if h == 1:
        return np.sin(th) / dx
    if h == 2:
        return (8 * np.sin(th) - np.sin(2 * th)) / (6 * dx)
    if h == 3:
        return (45 * np.sin(th) - 9 * np.sin(2 * th) + np.sin(3 * th)) / (30 * dx)
    if h == 4:
        return (672 * np.sin(th) - 168 * np.sin(2 * th)
                + 32 * np.sin(3 * th) - 3 * np.sin(4 * th)) / (420 * dx)
    return k


def second_eigenvalue(h, k, dx):
    """pystella/derivs.py:175-191 (SecondCenteredDifference.get_eigenvalues)."""
    th = k * dx
    if h == 1:
        return (2 * np.cos(th) - 2) / dx**2
    if h == 2:
        return (32 * np.cos(th) - 2 * np.cos(2 * th) - 30) / (12 * dx**2)
    if h == 3:
        return (90 * np.cos(th) - 9 * np.cos(2 * th)
                + 2/3 * np.cos(3 * th) - 245/3) / (30 * dx**2)
    if h == 4:
        return (1344 * np.cos(th) - 168 * np.cos(2 * th)
                + 64/3 * np.cos(3 * th) - 3/2 * np.cos(4 * th) - 7175/6
                ) / (420 * dx**2)
    return - k**2


def get_errs(a, b):
    """The reference's error metric, test/common.py:30-38."""
    a = np.atleast_1d(a)
    b = np.atleast_1d(b)
    mask = (a != 0.) | (b != 0.)
    a = a[mask]
    b = b[mask]
    if a.size == 0:
        return 0., 0.
    err = np.abs((a - b) / np.maximum(np.abs(a), np.abs(b)))
    return np.max(err), np.average(err)


# --------------------------------------------------------------------------
# halos -- pystella/decomp.py:177-182, 204-209, 232-233, 387-449
# --------------------------------------------------------------------------

def _h3(h):
    return (h,)*3 if isinstance(h, (int, np.integer)) else tuple(h)


def interior(arr, h):
    """View of the un-padded part of the last three axes."""
    h = _h3(h)
    slc = tuple(slice(hi, arr.shape[-3 + ax] - hi) for ax, hi in enumerate(h))
    return arr[(Ellipsis,) + slc]


def share_halos(arr, h):
    """
    Single-rank periodic halo fill, in place, order z -> x -> y, each pass
    spanning the *padded* extent of the other two axes so that edges and
    corners come out right (decomp.py:387-395, 422-424 with the
    pack_unpack_* statements at :181-182, :208-209, :232-233).
    """
    hx, hy, hz = _h3(h)
    Nx, Ny, Nz = (n - 2*hh for n, hh in zip(arr.shape[-3:], (hx, hy, hz)))
    if hz > 0:
        arr[..., :, :, 0:hz] = arr[..., :, :, Nz:Nz+hz]
        arr[..., :, :, Nz+hz:] = arr[..., :, :, hz:2*hz]
    if hx > 0:
        arr[..., 0:hx, :, :] = arr[..., Nx:Nx+hx, :, :]
        arr[..., Nx+hx:, :, :] = arr[..., hx:2*hx, :, :]
    if hy > 0:
        arr[..., :, 0:hy, :] = arr[..., :, Ny:Ny+hy, :]
        arr[..., :, Ny+hy:, :] = arr[..., :, hy:2*hy, :]
    return arr


# --------------------------------------------------------------------------
# stencils -- pystella/derivs.py:37-58, 61-108, 291-296
# --------------------------------------------------------------------------

def _shifted(fx, h, axis, s):
    """fx[i+h+s*(axis==0), j+h+s*(axis==1), k+h+s*(axis==2)] over the interior."""
    slc = []
    for ax in range(3):
        n = fx.shape[ax] - 2*h
        o = h + (s if ax == axis else 0)
        slc.append(slice(o, o + n))
    return fx[tuple(slc)]


def _second_stencil(fx, h, axis, coefs):
    # term order of centered_diff (derivs.py:94-108): for s in coefs: +s then -s
    out = None
    for s, c in coefs.items():
        term = c * _shifted(fx, h, axis, s)
        out = term if out is None else out + term
        if s != 0:
            out = out + c * _shifted(fx, h, axis, -s)
    return out


def _first_stencil(fx, h, axis, coefs):
    out = None
    for s, c in coefs.items():
        term = c * _shifted(fx, h, axis, s)
        out = term if out is None else out + term
        out = out + (-1 * c) * _shifted(fx, h, axis, -s)
    return out


def laplacian(fx, h, dx, halo=None):
    """
    ``lap = sum_axis stencil2(fx, axis) * dx_axis**-2`` (derivs.py:295-296).
    ``fx`` is padded by ``halo`` (defaults to ``h``) and must hold valid halos.
    """
    halo = h if halo is None else halo
    if halo != h:
        fx = fx[tuple(slice(halo - h, n - (halo - h)) for n in fx.shape)]
    coefs = LAP_COEFS[h]
    out = None
    for axis in range(3):
        term = _second_stencil(fx, h, axis, coefs) * dx[axis]**-2
        out = term if out is None else out + term
    return out


def partial(fx, h, dx, axis):
    """``pd_axis = stencil1(fx, axis) * (1/dx_axis)`` (derivs.py:293-294)."""
    return _first_stencil(fx, h, axis, GRAD_COEFS[h]) * (1 / dx[axis])


def gradient(fx, h, dx):
    return np.stack([partial(fx, h, dx, ax) for ax in range(3)])


def divergence(vec, h, dx):
    """pdx(v0) then += pdy(v1), += pdz(v2) (derivs.py:462-468)."""
    div = partial(vec[0], h, dx, 0)
    div = div + partial(vec[1], h, dx, 1)
    div = div + partial(vec[2], h, dx, 2)
    return div


# --------------------------------------------------------------------------
# time steppers -- pystella/step.py
# --------------------------------------------------------------------------

def _lsrk3ssp():
    # step.py:809-846
    c2 = .924574
    z1 = np.sqrt(36 * c2**4 + 36 * c2**3 - 135 * c2**2 + 84 * c2 - 12)
    z2 = 2 * c2**2 + c2 - 2
    z3 = 12 * c2**4 - 18 * c2**3 + 18 * c2**2 - 11 * c2 + 2
    z4 = 36 * c2**4 - 36 * c2**3 + 13 * c2**2 - 8 * c2 + 4
    z5 = 69 * c2**3 - 62 * c2**2 + 28 * c2 - 8
    z6 = 34 * c2**4 - 46 * c2**3 + 34 * c2**2 - 13 * c2 + 2
    B1 = c2
    B2 = ((12 * c2 * (c2 - 1) * (3 * z2 - z1) - (3 * z2 - z1)**2)
          / (144 * c2 * (3 * c2 - 2) * (c2 - 1)**2))
    B3 = (- 24 * (3 * c2 - 2) * (c2 - 1)**2
          / ((3 * z2 - z1)**2 - 12 * c2 * (c2 - 1) * (3 * z2 - z1)))
    A2 = ((- z1 * (6 * c2**2 - 4 * c2 + 1) + 3 * z3)
          / ((2 * c2 + 1) * z1 - 3 * (c2 + 2) * (2 * c2 - 1)**2))
    A3 = ((- z4 * z1 + 108 * (2 * c2 - 1) * c2**5 - 3 * (2 * c2 - 1) * z5)
          / (24 * z1 * c2 * (c2 - 1)**4 + 72 * c2 * z6
             + 72 * c2**6 * (2 * c2 - 13)))
    return [0, A2, A3], [B1, B2, B3]


#: name -> (A, B, expected_order); step.py:543-557 (RK54), :757-806 (RK3 family)
LSRK_TABLES = {
    "LowStorageRK54": (
        [0,
         -567301805773 / 1357537059087,
         -2404267990393 / 2016746695238,
         -3550918686646 / 2091501179385,
         -1275806237668 / 842570457699],
        [1432997174477 / 9575080441755,
         5161836677717 / 13612068292357,
         1720146321549 / 2090206949498,
         3134564353537 / 4481467310338,
         2277821191437 / 14882151754819],
        4),
    "LowStorageRK3Williamson": ([0, -5/9, -153/128], [1/3, 15/16, 8/15], 3),
    "LowStorageRK3Inhomogeneous": ([0, -17/32, -32/27], [1/4, 8/9, 3/4], 3),
    "LowStorageRK3Symmetric": ([0, -2/3, -1], [1/3, 1, 1/2], 3),
    "LowStorageRK3PredictorCorrector": ([0, -1/4, -4/3], [1/2, 2/3, 1/2], 3),
    "LowStorageRK3SSP": (*_lsrk3ssp(), 3),
}


def _long_tables():
    """
    The 12/13/14-stage schemes (step.py:568-754) are long literal tables; the
    oracle reads them from the committed extraction of the reference sources
    (tests/golden/reference_tables.json, see tests/golden/make_golden.py).
    """
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "tests", "golden", "reference_tables.json")
    if not os.path.exists(path):
        return
    with open(path) as fh:
        tabs = json.load(fh)["lsrk"]
    for name in ("LowStorageRK144", "LowStorageRK134", "LowStorageRK124"):
        LSRK_TABLES[name] = (tabs[name]["A"], tabs[name]["B"],
                             tabs[name]["expected_order"])


_long_tables()


def lsrk_stage(A_s, B_s, dt, unknowns, tmps, rhs):
    """
    One 2N-storage stage (step.py:476-485; statement order elementwise.py:173-203):
    all right-hand sides are evaluated from the *old* state first, then for
    each unknown ``k <- A*k + dt*rhs ; u <- u + B*k``.  ``unknowns``/``tmps`` are
    lists of arrays updated in place (only where ``rhs`` is defined, i.e. the
    caller passes interior views); ``rhs`` is a list of arrays.
    """
    for u, k, r in zip(unknowns, tmps, rhs):
        k[...] = A_s * k + dt * r
        u[...] = u + B_s * k


#: classical RK: name -> (num_stages, num_copies, expected_order)
RK_TABLES = {
    "RungeKutta4": (4, 3, 4),
    "RungeKutta3Heun": (3, 3, 3),
    "RungeKutta3Nystrom": (3, 3, 3),
    "RungeKutta3Ralston": (3, 3, 3),
    "RungeKutta3SSP": (3, 2, 3),
    "RungeKutta2Midpoint": (2, 2, 2),
    "RungeKutta2Heun": (2, 2, 2),
    "RungeKutta2Ralston": (2, 2, 2),
}


def rk_rhs_copy(stage):
    """RHS reads copy q=0 at stage 0, else q=1 (step.py:231)."""
    return 0 if stage == 0 else 1


def rk_stage(name, stage, dt, fq, rhs):
    """
    Classical RK stage on an array ``fq`` with leading copies axis
    (step.py:250-411: the ``step_statements`` of each subclass).
    ``rhs`` was evaluated from copy ``rk_rhs_copy(stage)``.
    """
    f = fq
    if name == "RungeKutta4":
        if stage == 0:
            f[1], f[2] = f[0] + dt/2 * rhs, f[0] + dt/6 * rhs
        elif stage == 1:
            f[1], f[2] = f[0] + dt/2 * rhs, f[2] + dt/3 * rhs
        elif stage == 2:
            f[1], f[2] = f[0] + dt * rhs, f[2] + dt/3 * rhs
        elif stage == 3:
            f[0] = f[2] + dt/6 * rhs
    elif name == "RungeKutta3Heun":
        if stage == 0:
            f[1], f[2] = f[0] + dt/3 * rhs, f[0] + dt/4 * rhs
        elif stage == 1:
            f[1] = f[0] + dt*2/3 * rhs
        elif stage == 2:
            f[0] = f[2] + dt*3/4 * rhs
    elif name == "RungeKutta3Nystrom":
        if stage == 0:
            f[1], f[2] = f[0] + dt*2/3 * rhs, f[0] + dt*2/8 * rhs
        elif stage == 1:
            f[1], f[2] = f[0] + dt*2/3 * rhs, f[2] + dt*3/8 * rhs
        elif stage == 2:
            f[0] = f[2] + dt*3/8 * rhs
    elif name == "RungeKutta3Ralston":
        if stage == 0:
            f[1], f[2] = f[0] + dt/2 * rhs, f[0] + dt*2/9 * rhs
        elif stage == 1:
            f[1], f[2] = f[0] + dt*3/4 * rhs, f[2] + dt*1/3 * rhs
        elif stage == 2:
            f[0] = f[2] + dt*4/9 * rhs
    elif name == "RungeKutta3SSP":
        if stage == 0:
            f[1] = f[0] + dt * rhs
        elif stage == 1:
            f[1] = 3/4 * f[0] + 1/4 * f[1] + dt/4 * rhs
        elif stage == 2:
            f[0] = 1/3 * f[0] + 2/3 * f[1] + dt*2/3 * rhs
    elif name == "RungeKutta2Midpoint":
        if stage == 0:
            f[1] = f[0] + dt/2 * rhs
        elif stage == 1:
            f[0] = f[0] + dt * rhs
    elif name == "RungeKutta2Heun":
        if stage == 0:
            f[1], f[0] = f[0] + dt * rhs, f[0] + dt/2 * rhs
        elif stage == 1:
            f[0] = f[0] + dt/2 * rhs
    elif name == "RungeKutta2Ralston":
        if stage == 0:
            f[1], f[0] = f[0] + dt*2/3 * rhs, f[0] + dt/4 * rhs
        elif stage == 1:
            f[0] = f[0] + dt*3/4 * rhs
    else:
        raise NotImplementedError(name)


# --------------------------------------------------------------------------
# sectors -- pystella/sectors.py
# --------------------------------------------------------------------------

def poly_eval(terms, fields):
    """
    Evaluate a polynomial given as ``[(coef, (p_0, ..., p_{n-1})), ...]`` on a
    list of field arrays.  (The potentials of examples/scalar_preheating.py:98-105
    and test/test_energy.py:49-51 are of this form.)  Integer powers are expanded
    into repeated products, ``coef * f_0 * f_0 * ... * f_1 * ...``, summed left to
    right -- what a code generator emits for small integer exponents, and the
    association order the CUDA kernels use.
    """
    out = 0.
    for c, powers in terms:
        t = c
        for fld, p in zip(fields, powers):
            for _ in range(int(p)):
                t = t * fld
        out = out + t
    return out


def poly_diff(terms, n):
    """d/df_n of a polynomial in the representation of :func:`poly_eval`."""
    out = []
    for c, powers in terms:
        p = powers[n]
        if p == 0:
            continue
        new = list(powers)
        new[n] = p - 1
        out.append((c * p, tuple(new)))
    return out


def scalar_rhs(f, dfdt, lap_f, a, hubble, potential_terms):
    """
    ScalarSector.rhs_dict (sectors.py:117-131): for each field
    ``f' = dfdt`` ; ``dfdt' = lap_f - 2*hubble*dfdt - a**2 * dV/df``.
    ``f``, ``dfdt``, ``lap_f`` are lists/arrays of *interior* arrays.
    Returns the RHS list in rhs_dict order ``[f0, dfdt0, f1, dfdt1, ...]``.
    """
    n = len(f)
    rhs = []
    for fld in range(n):
        dV = poly_eval(poly_diff(potential_terms, fld), f) if potential_terms else 0.
        rhs.append(dfdt[fld])
        rhs.append(lap_f[fld] - 2 * hubble * dfdt[fld] - a**2 * dV)
    return rhs


def tensor_index(i, j):
    """sectors.py:164-167."""
    a = i if i <= j else j
    b = j if i <= j else i
    return (7 - a) * a // 2 - 4 + b


def gw_rhs(dhijdt, lap_hij, dfdx, hubble):
    """
    TensorPerturbationSector.rhs_dict (sectors.py:187-204) with the ScalarSector
    stress tensor, drop_trace=True (sectors.py:146-152):
    ``hij' = dhijdt`` ; ``dhijdt' = lap_hij - 2*hubble*dhijdt
    + 16*pi * sum_fld dfdx[fld, i-1]*dfdx[fld, j-1]``.
    ``dfdx`` has shape ``(nscalars, 3, Nx, Ny, Nz)``.  Returns RHS in rhs_dict order.
    """
    rhs = []
    for i in range(1, 4):
        for j in range(i, 4):
            fld = tensor_index(i, j)
            Sij = sum(dfdx[n][i-1] * dfdx[n][j-1] for n in range(len(dfdx)))
            rhs.append(dhijdt[fld])
            rhs.append(lap_hij[fld] - 2 * hubble * dhijdt[fld] + 16 * np.pi * Sij)
    return rhs


# --------------------------------------------------------------------------
# reductions -- pystella/reduction.py:141-173, 237-255; sectors.py:133-144, 211-229
# --------------------------------------------------------------------------

def _sum_x_first(arr):
    # reduction.py:165-173: tmp[j,k] = sum_i expr ; then a tree sum of tmp
    return np.sum(np.sum(arr, axis=0))


def scalar_energy(f, dfdt, lap_f, a, potential_terms, grid_size=None):
    """
    ScalarSector.reducers averaged over the grid: kinetic[n] = <dfdt_n**2/2/a**2>,
    potential = <V(f)>, gradient[n] = <-f_n*lap_f_n/2/a**2> (sectors.py:139-143).
    Inputs are interior arrays (lists over fields).
    """
    n = len(f)
    gs = grid_size or f[0].size
    out = {}
    out["kinetic"] = np.array(
        [_sum_x_first(dfdt[i]**2 / 2 / a**2) / gs for i in range(n)])
    V = poly_eval(potential_terms, f) if potential_terms else 0. * f[0]
    out["potential"] = np.array([_sum_x_first(V + 0. * f[0]) / gs])
    out["gradient"] = np.array(
        [_sum_x_first(- f[i] * lap_f[i] / 2 / a**2) / gs for i in range(n)])
    return out


def get_rho_and_p(energy):
    """sectors.py:211-229."""
    energy["total"] = sum(sum(e) for e in energy.values())
    energy["pressure"] = 0
    if "kinetic" in energy:
        energy["pressure"] += sum(energy["kinetic"])
    if "gradient" in energy:
        energy["pressure"] += - sum(energy["gradient"]) / 3
    if "potential" in energy:
        energy["pressure"] += - sum(energy["potential"])
    return energy


def field_statistics(f, h, max_min=False):
    """FieldStatistics.__call__ (reduction.py:284-343): loops outer axes."""
    outer = f.shape[:-3]
    out = {k: np.zeros(outer) for k in ("mean", "variance")}
    if max_min:
        for k in ("max", "min", "abs_max", "abs_min"):
            out[k] = np.zeros(outer)
    for s in np.ndindex(*outer):
        x = interior(f[s], h)
        mean = _sum_x_first(x) / x.size
        out["mean"][s] = mean
        out["variance"][s] = _sum_x_first(x**2) / x.size - mean**2
        if max_min:
            out["max"][s] = x.max()
            out["min"][s] = x.min()
            out["abs_max"][s] = np.abs(x).max()
            out["abs_min"][s] = np.abs(x).min()
    return out


# --------------------------------------------------------------------------
# expansion -- pystella/expansion.py:72-157, 176
# --------------------------------------------------------------------------

class Expansion:
    """
    FLRW background stepped with the same (low-storage) RK coefficients as the
    fields (expansion.py:72-99, 140-157).  Only LowStorageRK* steppers here.
    """

    def __init__(self, energy, stepper="LowStorageRK54", mpl=1., dtype=np.float64):
        self.mpl = mpl
        self.A, self.B, _ = LSRK_TABLES[stepper]
        self.a = np.ones((1,), dtype=dtype)
        self.adot = self.adot_friedmann_1(self.a, energy)
        self.hubble = self.adot / self.a
        self._ka = np.zeros((1,), dtype=dtype)
        self._kadot = np.zeros((1,), dtype=dtype)

    def adot_friedmann_1(self, a, energy):
        """expansion.py:101-117."""
        return np.sqrt(8 * np.pi * a**2 / 3 / self.mpl**2 * energy) * a

    def addot_friedmann_2(self, a, energy, pressure):
        """expansion.py:119-138."""
        return 4 * np.pi * a**2 / 3 / self.mpl**2 * (energy - 3 * pressure) * a

    def step(self, stage, energy, pressure, dt):
        """expansion.py:140-157 with the LSRK stage of step.py:476-485."""
        rhs_a = self.adot.copy()
        rhs_adot = self.addot_friedmann_2(self.a, energy, pressure)
        lsrk_stage(self.A[stage], self.B[stage], dt,
                   [self.a, self.adot], [self._ka, self._kadot],
                   [rhs_a, rhs_adot])
        self.hubble[()] = self.adot / self.a

    def constraint(self, energy):
        """expansion.py:159-176."""
        return np.abs(self.adot_friedmann_1(self.a[0], energy) / self.adot[0] - 1)


# --------------------------------------------------------------------------
# domain decomposition arithmetic -- pystella/decomp.py:137-139, 287-337
# --------------------------------------------------------------------------

def rank_id(proc_shape, rx, ry, rz):
    """decomp.py:287-296."""
    rx_ = (rx + proc_shape[0]) % proc_shape[0]
    ry_ = (ry + proc_shape[1]) % proc_shape[1]
    rz_ = (rz + proc_shape[2]) % proc_shape[2]
    return rz_ + proc_shape[2] * (ry_ + proc_shape[1] * rx_)


def rank_tuple_of(proc_shape, rank):
    """decomp.py:137-139."""
    rz = rank % proc_shape[2]
    ry = (rank - rz) // proc_shape[2] % proc_shape[1]
    rx = (rank - rz - proc_shape[2] * ry) // proc_shape[1]
    return rx, ry, rz


def rank_shape_start(proc_shape, grid_shape, rank_tuple):
    """decomp.py:322-337 (matches mpi4py_fft.pencil)."""
    shape, start = [], []
    for N, P, r in zip(grid_shape, proc_shape, rank_tuple):
        q, rem = divmod(N, P)
        shape.append(q + (1 if rem > r else 0))
        start.append(r * q + min(r, rem))
    return tuple(shape), tuple(start)


# --------------------------------------------------------------------------
# the reference's unfused per-stage schedule (examples/scalar_preheating.py:258-266,
# examples/wave_equation.py:61-65) for a ScalarSector, single rank
# --------------------------------------------------------------------------

def reference_schedule_stage(stage, f, dfdt, kf, kdfdt, lap_f, h, dx, dt,
                             a=1., hubble=0., potential_terms=None,
                             tables="LowStorageRK54"):
    """
    One stage as the reference's user loop performs it for ``lap_f`` computed
    from the *current* ``f`` (``derivs`` -> ``stepper``):
    share_halos(f) ; lap_f = lap(f) ; LSRK stage on (f, dfdt).
    ``f``, ``dfdt``, ``kf``, ``kdfdt`` have shape ``(n, padded)``, ``lap_f``
    ``(n, Nx, Ny, Nz)``; all updated in place.
    """
    A, B, _ = LSRK_TABLES[tables]
    n = f.shape[0]
    for fld in range(n):
        share_halos(f[fld], h)
        lap_f[fld] = laplacian(f[fld], h, dx)
    fi = [interior(f[i], h) for i in range(n)]
    di = [interior(dfdt[i], h) for i in range(n)]
    rhs = scalar_rhs(fi, di, [lap_f[i] for i in range(n)], a, hubble,
                     potential_terms)
    # materialise RHS before any update (tmp_instructions come first)
    rhs = [np.array(r, copy=True) for r in rhs]
    unknowns, tmps = [], []
    for i in range(n):
        unknowns += [fi[i], di[i]]
        tmps += [interior(kf[i], h), interior(kdfdt[i], h)]
    lsrk_stage(A[stage], B[stage], dt, unknowns, tmps, rhs)
