"""
Generates ``tests/golden/reference_tables.json`` from the *reference sources*.

The reference (zachjweiner/pystella) is pure Python but imports loopy /
pyopencl / pymbolic / mpi4py at module scope, none of which exist in this
image.  Its numeric tables (finite-difference coefficients, stencil
eigenvalues, Runge-Kutta coefficients, tensor_index, rank partitioning) are
plain Python, though, so this script installs permissive *stub* modules for
the missing third-party packages, imports the unmodified reference modules
from ``/root/reference`` and dumps those tables.  Only run in the build
container (the reference is not present on the GPU box); the JSON it writes is
committed and is what ``tests/test_oracle_golden.py`` checks the oracle and the
product tables against.

    python tests/golden/make_golden.py
"""

import importlib.abc
import importlib.machinery
import json
import os
import sys
import types

REFERENCE = os.environ.get("PYSTELLA_REFERENCE", "/root/reference")
STUBBED = ("pymbolic", "loopy", "pyopencl", "mpi4py", "pytools", "h5py",
           "pycl_fft", "mpi4py_fft", "islpy", "gpyfft", "pyvkfft")


class _Meta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy

    def __getitem__(cls, item):
        return _Dummy


class _Dummy(metaclass=_Meta):
    """Usable as a base class, a callable, a decorator and an attribute bag."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def main():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REFERENCE)
    import numpy as np
    import pystella.derivs as derivs
    import pystella.step as step
    from pystella.sectors import tensor_index

    out = {"source": "zachjweiner/pystella @ /root/reference (unmodified modules, "
                     "third-party imports stubbed)"}

    out["grad_coefs"] = {str(h): {str(s): c for s, c in v.items()}
                         for h, v in derivs._grad_coefs.items()}
    out["lap_coefs"] = {str(h): {str(s): c for s, c in v.items()}
                        for h, v in derivs._lap_coefs.items()}

    # stencil eigenvalues at the test's wavenumbers (test/test_derivs.py:49-62)
    L = (3, 5, 7)
    grid = (32, 32, 32)
    dx = [Li / Ni for Li, Ni in zip(L, grid)]
    kvec = [2 * np.pi / Li * m for Li, m in zip(L, (-5, 4, -3))]
    eig = {}
    for h in (1, 2, 3, 4):
        e1 = derivs.FirstCenteredDifference(h)
        e2 = derivs.SecondCenteredDifference(h)
        eig[str(h)] = {
            "first": [float(e1.get_eigenvalues(k, d)) for k, d in zip(kvec, dx)],
            "second": [float(e2.get_eigenvalues(k, d)) for k, d in zip(kvec, dx)],
        }
    out["eigenvalues"] = {"L": L, "grid": grid, "modes": [-5, 4, -3], "values": eig}

    lsrk = {}
    rk = {}
    for cls in vars(step).values():
        if not isinstance(cls, type):
            continue
        if issubclass(cls, step.LowStorageRKStepper) and cls._A:
            lsrk[cls.__name__] = {
                "A": [float(x) for x in cls._A], "B": [float(x) for x in cls._B],
                "C": [float(x) for x in cls._C],
                "num_stages": cls.num_stages, "expected_order": cls.expected_order}
        elif (issubclass(cls, step.RungeKuttaStepper)
              and cls.num_stages is not None):
            rk[cls.__name__] = {
                "num_stages": cls.num_stages, "num_copies": cls.num_copies,
                "expected_order": cls.expected_order}
    out["lsrk"] = lsrk
    out["rk"] = rk
    out["all_steppers"] = [c.__name__ for c in step.all_steppers]

    out["tensor_index"] = {f"{i},{j}": tensor_index(i, j)
                           for i in range(1, 4) for j in range(1, 4)}

    # rank partitioning (decomp.py:322-337) without constructing the class
    import pystella.decomp as decomp

    class _D:
        pass

    parts = []
    for proc_shape, grid_shape in [((2, 2, 1), (32, 32, 32)), ((3, 1, 1), (109, 67, 32)),
                                   ((8, 1, 1), (512, 512, 512)), ((4, 2, 1), (130, 67, 9))]:
        d = _D()
        d.proc_shape = proc_shape
        for rx in range(proc_shape[0]):
            for ry in range(proc_shape[1]):
                d.rank_tuple = (rx, ry, 0)
                shape, start = decomp.DomainDecomposition.get_rank_shape_start(
                    d, grid_shape)
                rid = decomp.DomainDecomposition.rankID(d, rx, ry, 0)
                rid_m = decomp.DomainDecomposition.rankID(d, rx - 1, ry - 1, 0)
                parts.append({"proc_shape": proc_shape, "grid_shape": grid_shape,
                              "rank_tuple": (rx, ry, 0), "rank_shape": shape,
                              "rank_start": start, "rankID": rid,
                              "rankID_minus": rid_m})
    out["partition"] = parts

    # Fourier-space helpers of the initial conditions (pure NumPy in the reference)
    import pystella.fourier.dft as rdft
    import pystella.fourier.rayleigh as rray

    class _Arr:                      # stands in for a pyopencl array: only .get() is used
        def __init__(self, a):
            self.a = a

        def get(self):
            return self.a

    out["fftfreq"] = {str(n): [float(x) for x in rdft.fftfreq(n)] for n in (4, 7, 8, 32)}
    fgrid = (8, 6, 4)
    kshape = (8, 6, 3)
    rng = np.random.default_rng(20240917)
    fk = rng.standard_normal(kshape) + 1j * rng.standard_normal(kshape)
    herm = rray.make_hermitian(fk.copy())

    class _F:
        grid_shape = fgrid
        sub_k = {n: _Arr(k) for n, k in zip(
            ("momenta_x", "momenta_y", "momenta_z"),
            [rdft.fftfreq(fgrid[0]), rdft.fftfreq(fgrid[1]),
             np.fft.rfftfreq(fgrid[2], 1 / fgrid[2])])}

    zeroed_imag = rdft.BaseDFT.zero_corner_modes(_F(), herm.copy(), only_imag=True)
    zeroed = rdft.BaseDFT.zero_corner_modes(_F(), fk.copy())
    wk = np.sqrt(rng.random(kshape) + 1.0)
    km = wk.copy()
    km[0, 0, 0] = 0.
    wk_in = wk.copy()
    power = rray.RayleighGenerator._ps_wrapper(None, lambda w: 1 / 2 / w, wk_in, km)
    out["fourier"] = {
        "grid_shape": fgrid, "seed": 20240917,
        "fk": [fk.real.tolist(), fk.imag.tolist()],
        "make_hermitian": [herm.real.tolist(), herm.imag.tolist()],
        "zero_corner_modes_only_imag": [zeroed_imag.real.tolist(), zeroed_imag.imag.tolist()],
        "zero_corner_modes": [zeroed.real.tolist(), zeroed.imag.tolist()],
        "ps_wrapper": {"wk": wk.tolist(), "power": power.tolist(),
                       "wk_after": wk_in.tolist()},
    }

    # the one stored golden number: test/test_examples.py:31-34, and the defaults of the
    # example it runs (examples/scalar_preheating.py:32-66, 80-82)
    import re
    text = open(os.path.join(REFERENCE, "test", "test_examples.py")).read()
    m = re.search(r'"examples/scalar_preheating.py":\s*([0-9.eE+-]+)', text)
    ex = open(os.path.join(REFERENCE, "examples", "scalar_preheating.py")).read()

    def default(flag):
        mm = re.search(r'add_argument\("--' + flag + r'".*?default=(\([^)]*\)|[^,\n]+),', ex, re.S)
        return mm.group(1).strip()

    out["example"] = {
        "constraint": float(m.group(1)),
        "argv": ["--grid-shape", "32", "32", "32", "--end-time", "1"],
        "defaults": {k: default(k) for k in ("halo-shape", "box-dim", "kappa", "mpl", "mphi",
                                             "mchi", "gsq", "sigma", "lambda4")},
        "f0": re.search(r"f0 = \[(.*?)\]", ex).group(1),
        "df0": re.search(r"df0 = \[(.*?)\]", ex).group(1),
    }

    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "reference_tables.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote", os.path.join(here, "reference_tables.json"))


if __name__ == "__main__":
    main()
