"""TEST INFRASTRUCTURE ONLY -- ctypes access to oracle_c.c (checker / CPU baseline)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAXS, MAXT = 4, 16


class Potential(C.Structure):
    _fields_ = [("nscalars", C.c_int), ("nterms", C.c_int * MAXS),
                ("coef", (C.c_double * MAXT) * MAXS),
                ("power", ((C.c_int * MAXS) * MAXT) * MAXS),
                ("nvterms", C.c_int), ("vcoef", C.c_double * MAXT),
                ("vpower", (C.c_int * MAXS) * MAXT)]


def build(native=False):
    target = "native" if native else "all"
    subprocess.check_call(["make", "-s", "-C", HERE, target])
    name = "liboracle_c_native.so" if native else "liboracle_c.so"
    return os.path.join(HERE, "_build", name)


_lib = {}


def load(native=False):
    if native not in _lib:
        path = os.path.join(HERE, "_build",
                            "liboracle_c_native.so" if native else "liboracle_c.so")
        if native or not os.path.exists(path):
            try:
                path = build(native)
            except Exception:
                if native:
                    return load(False)
                raise
        lib = C.CDLL(path)
        lib.orc_num_threads.restype = C.c_int
        _lib[native] = lib
    return _lib[native]


def make_potential(nscalars, V_terms):
    from oracle.numpy_oracle import poly_diff
    p = Potential()
    p.nscalars = nscalars
    V_terms = V_terms or []
    for n in range(nscalars):
        d = poly_diff(V_terms, n)
        p.nterms[n] = len(d)
        for t, (c, pw) in enumerate(d):
            p.coef[n][t] = c
            for m in range(nscalars):
                p.power[n][t][m] = pw[m]
    p.nvterms = len(V_terms)
    for t, (c, pw) in enumerate(V_terms):
        p.vcoef[t] = c
        for m in range(nscalars):
            p.vpower[t][m] = pw[m]
    return p


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def share_halos(a, h, native=False):
    Nx, Ny, Nz = (n - 2 * h for n in a.shape)
    load(native).orc_share_halos(_p(a), C.c_int64(Nx), C.c_int64(Ny), C.c_int64(Nz), h)


def laplacian(f, lap, h, dx, native=False):
    Nx, Ny, Nz = lap.shape
    dxa = np.asarray(dx, dtype=np.float64)
    load(native).orc_laplacian(_p(f), _p(lap), C.c_int64(Nx), C.c_int64(Ny), C.c_int64(Nz),
                               h, _p(dxa))


def scalar_energy(f, dfdt, lap, h, a, grid_size, pot, native=False):
    Nx, Ny, Nz = lap.shape[-3:]
    out = np.zeros(2 * pot.nscalars + 1)
    load(native).orc_scalar_energy(_p(f), _p(dfdt), _p(lap), C.c_int64(Nx), C.c_int64(Ny),
                                   C.c_int64(Nz), h, C.c_double(a), C.c_double(grid_size),
                                   C.byref(pot), _p(out))
    return out


def preheating_steps(nsteps, A, B, f, dfdt, kf, kd, lap, h, dx, dt, mpl, grid_size, pot,
                     state, energies, with_expansion=True, native=False):
    Nx, Ny, Nz = lap.shape[-3:]
    Aa, Ba = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
    dxa = np.asarray(dx, dtype=np.float64)
    load(native).orc_preheating_steps(
        nsteps, len(Aa), _p(Aa), _p(Ba), _p(f), _p(dfdt), _p(kf), _p(kd), _p(lap),
        C.c_int64(Nx), C.c_int64(Ny), C.c_int64(Nz), h, _p(dxa), C.c_double(dt),
        C.c_double(mpl), C.c_double(grid_size), C.byref(pot), _p(state), _p(energies),
        int(with_expansion))


def gradient(f, grd, h, dx, native=False):
    Nx, Ny, Nz = grd.shape[-3:]
    dxa = np.asarray(dx, dtype=np.float64)
    load(native).orc_gradient(_p(f), _p(grd), C.c_int64(Nx), C.c_int64(Ny), C.c_int64(Nz), h,
                              _p(dxa))


def tensor_stage(hij, dhij, kh, kdh, lap_hij, dfdx, h, dt, A, B, hubble, native=False):
    Nx, Ny, Nz = lap_hij.shape[-3:]
    load(native).orc_lsrk_tensor_stage(
        _p(hij), _p(dhij), _p(kh), _p(kdh), _p(lap_hij), _p(dfdx), int(dfdx.shape[0]),
        C.c_int64(Nx), C.c_int64(Ny), C.c_int64(Nz), h, C.c_double(dt), C.c_double(A),
        C.c_double(B), C.c_double(hubble))


def preheating_gw_steps(nsteps, A, B, f, dfdt, kf, kd, lap_f, dfdx, hij, dhij, kh, kdh, lap_hij,
                        h, dx, dt, mpl, grid_size, pot, state, energies, with_expansion=True,
                        native=False):
    """Scalars + tensor perturbations + background, loop of scalar_preheating.py:258-266."""
    Nx, Ny, Nz = lap_f.shape[-3:]
    Aa, Ba = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
    dxa = np.asarray(dx, dtype=np.float64)
    load(native).orc_preheating_gw_steps(
        nsteps, len(Aa), _p(Aa), _p(Ba), _p(f), _p(dfdt), _p(kf), _p(kd), _p(lap_f), _p(dfdx),
        _p(hij), _p(dhij), _p(kh), _p(kdh), _p(lap_hij),
        C.c_int64(Nx), C.c_int64(Ny), C.c_int64(Nz), h, _p(dxa), C.c_double(dt),
        C.c_double(mpl), C.c_double(grid_size), C.byref(pot), _p(state), _p(energies),
        int(with_expansion))


def gw_initial_derivatives(f, hij, h, dx, native=False):
    """share_halos + (lap_f, dfdx, lap_hij) of the initial state, as the example computes
    them before its loop (scalar_preheating.py:236-247)."""
    ns = f.shape[0]
    N = tuple(n - 2 * h for n in f.shape[-3:])
    lap_f, dfdx, lap_hij = np.zeros((ns,) + N), np.zeros((ns, 3) + N), np.zeros((6,) + N)
    for n in range(ns):
        share_halos(f[n], h, native)
        laplacian(f[n], lap_f[n], h, dx, native)
        gradient(f[n], dfdx[n], h, dx, native)
    for c in range(6):
        share_halos(hij[c], h, native)
        laplacian(hij[c], lap_hij[c], h, dx, native)
    return lap_f, dfdx, lap_hij


def scalar_stage(f, dfdt, kf, kd, lap, h, dt, A, B, a, hubble, pot, native=False):
    """One LSRK stage of the scalar sector (orc_lsrk_scalar_stage)."""
    Nx, Ny, Nz = lap.shape[-3:]
    load(native).orc_lsrk_scalar_stage(
        _p(f), _p(dfdt), _p(kf), _p(kd), _p(lap), C.c_int64(Nx), C.c_int64(Ny),
        C.c_int64(Nz), h, C.c_double(dt), C.c_double(A), C.c_double(B), C.c_double(a),
        C.c_double(hubble), C.byref(pot))


def first_touch_copy(src, native=False):
    """A copy of float64 ``src`` whose pages are first touched by the OpenMP threads that will
    work on them (NUMA placement of the timed CPU baseline)."""
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.empty_like(src)
    row = src.shape[-1] if src.ndim else 1
    load(native).orc_first_touch_copy(_p(dst), _p(src), C.c_int64(src.size // row),
                                      C.c_int64(row))
    return dst


def first_touch_zeros(shape, native=False):
    dst = np.empty(shape, dtype=np.float64)
    row = dst.shape[-1]
    load(native).orc_first_touch_copy(_p(dst), None, C.c_int64(dst.size // row),
                                      C.c_int64(row))
    return dst


def num_threads(native=False):
    return load(native).orc_num_threads()
