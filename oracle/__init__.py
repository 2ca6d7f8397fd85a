"""
TEST INFRASTRUCTURE ONLY.

CPU restatement of the zachjweiner/pystella hot path (finite-difference
stencils, low-storage RK stages, periodic halo fill, reductions, FLRW
expansion).  Nothing under ``pystella_b200/`` may import this package: it is
the checker used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Parity status: the reference itself (pyopencl + loopy + pymbolic) cannot be
imported or run in this image, so the oracle is pinned by (i) the coefficient
tables extracted from the reference sources by ``tests/golden/make_golden.py``
(committed as ``tests/golden/reference_tables.json``) and (ii) the reference's
own known-answer tests re-expressed in ``tests/test_oracle_kat.py``
(plane-wave eigenvalues, ``y' = y**n`` convergence, bit-exact periodic halos,
FLRW expansion) and (iii) the reference's one stored golden number: the Friedmann
constraint of ``examples/scalar_preheating.py`` at 32^3 (``test/test_examples.py:33``),
which the oracle reproduces end to end -- WKB Rayleigh initial conditions
(``oracle/fourier_oracle.py``), stencils, LowStorageRK54, energies, expansion -- to 0.2%
(statistically: the reference's random stream is pyopencl's) and (iv) the right-hand sides,
reducers, dV/df, stress tensor and rho / P that the reference's own ``pystella/sectors.py``
and ``pystella/field`` produce on seeded random data (``tests/golden/make_golden_sectors.py``
imports them unmodified over a stand-in for pymbolic's expression nodes; fixture
``tests/golden/reference_sectors.json``, checked in ``tests/test_reference_sectors_golden.py``
to 2e-13).  Per-stage array parity against the generated pyopencl kernels themselves is
unpinned by the reference (it stores no array fixtures).
"""

from oracle.numpy_oracle import *  # noqa: F401,F403
