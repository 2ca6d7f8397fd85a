"""
FLRW background: :class:`Expansion` with the interface of
``pystella/expansion.py:28-176``.  Two host scalars (``a``, ``adot``) stepped
with the *same* stepper class as the fields; the reference compiles this tiny ODE
with loopy's C target (``expansion.py:91-99``), here the stepper's host path
evaluates it with NumPy.
"""

import numpy as np

from pystella_b200.field import Field, var

__all__ = ["Expansion"]


class Expansion:
    """
    :arg energy: initial energy density (initialises ``adot`` through the first
        Friedmann equation).
    :arg Stepper: the stepper *class* (e.g. :class:`~pystella_b200.LowStorageRK54`).
    :arg mpl: unreduced Planck mass (sets the units); default 1.
    :arg dtype: default ``float64``.
    """

    def __init__(self, energy, Stepper, mpl=1., dtype=np.float64):
        self.mpl = mpl
        from pystella_b200.step import LowStorageRKStepper
        self.is_low_storage = issubclass(Stepper, LowStorageRKStepper)
        num_copies = Stepper.num_copies or 1
        self.a = np.ones((num_copies,), dtype=dtype)
        self.adot = self.adot_friedmann_1(self.a, energy)
        self.hubble = self.adot / self.a

        if self.is_low_storage:
            a_ = Field("a", indices=[], shape=(1,))[0]
            adot_ = Field("adot", indices=[], shape=(1,))[0]
        else:
            a_ = Field("a", indices=[])
            adot_ = Field("adot", indices=[])
        rhs_dict = {a_: adot_,
                    adot_: self.addot_friedmann_2(a_, var("energy"), var("pressure"))}
        self.stepper = Stepper(rhs_dict, rank_shape=(0, 0, 0), halo_shape=0, dtype=dtype)

    def adot_friedmann_1(self, a, energy):
        """``a' = sqrt(8 pi a^2 rho / (3 mpl^2)) a`` (expansion.py:101-117)."""
        return np.sqrt(8 * np.pi * a**2 / 3 / self.mpl**2 * energy) * a

    def addot_friedmann_2(self, a, energy, pressure):
        """``a'' = 4 pi a^2 (rho - 3 P) / (3 mpl^2) a`` (expansion.py:119-138)."""
        return 4 * np.pi * a**2 / 3 / self.mpl**2 * (energy - 3 * pressure) * a

    def step(self, stage, energy, pressure, dt):
        """One stage of the background integrator (expansion.py:140-157)."""
        self.stepper(stage, a=self.a, adot=self.adot, dt=dt,
                     energy=_scalar(energy), pressure=_scalar(pressure))
        self.hubble[()] = self.adot / self.a

    def constraint(self, energy):
        """``|sqrt(8 pi a^2 rho/(3 mpl^2)) / H - 1|`` (expansion.py:159-176)."""
        return np.abs(self.adot_friedmann_1(self.a[0], energy) / self.adot[0] - 1)


def _scalar(x):
    arr = np.asarray(x)
    return arr.reshape(-1)[0] if arr.size == 1 else arr
