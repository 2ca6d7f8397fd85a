"""
FLRW background: :class:`Expansion` with the interface of
``pystella/expansion.py:28-176``.  Two host scalars (``a``, ``adot``) stepped
with the *same* stepper class as the fields; the reference compiles this tiny ODE
with loopy's C target (``expansion.py:91-99``), here the stepper's host path
evaluates it with NumPy.
"""

import numpy as np

from pystella_b200.field import Field, var

__all__ = ["Expansion", "DeviceExpansion"]


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


class DeviceExpansion:
    """
    The same background ODE with its state resident on the GPU, fed by the energy
    sums the fused stage kernel accumulates (``energy=True`` steppers), so that a
    time step needs no device-to-host synchronisation::

        expand = ps.DeviceExpansion(energy0, ps.LowStorageRK54, grid_size=N**3)
        for s in range(stepper.num_stages):
            stepper(s, f=f, dfdt=dfdt, a=expand.a, hubble=expand.hubble)
            expand.step(s, stepper, dt)

    Stage ``s`` of the stepper reads ``(a, hubble)`` and accumulates the energy sums
    of its *input* state; :meth:`step` turns them into energy density and pressure
    with the current ``a`` and advances ``(a, adot)`` -- the same values, in the same
    order, as the reference loop ``stepper -> expand.step -> compute_energy``
    (``examples/scalar_preheating.py:258-264``), which recomputes the Laplacian
    and synchronises with the host three times per stage to get them.

    Attributes ``a``, ``adot``, ``hubble`` are 1-element float64 device tensors
    (views of :attr:`state`); ``energies`` holds, after each :meth:`step`,
    ``{kinetic_n.., potential, gradient_n.., total, pressure}`` of the state that
    entered the stage.
    """

    def __init__(self, energy, Stepper, mpl=1., grid_size=None, decomp=None,
                 nscalars=1, a0=1., device=None, strict_fp=False):
        import torch
        from pystella_b200.step import LowStorageRKStepper
        if not issubclass(Stepper, LowStorageRKStepper):
            raise NotImplementedError("DeviceExpansion needs a low-storage stepper")
        if grid_size is None:
            raise ValueError("pass grid_size= (global number of lattice sites)")
        self.mpl = float(mpl)
        self.grid_size = float(grid_size)
        self.decomp = decomp
        self.nscalars = nscalars
        self.strict_fp = strict_fp
        self._A, self._B = Stepper._A, Stepper._B
        self.num_stages = Stepper.num_stages
        dev = device or torch.device("cuda", torch.cuda.current_device())
        adot0 = float(np.sqrt(8 * np.pi * a0**2 / 3 / self.mpl**2 * float(energy)) * a0)
        self.state = torch.tensor([a0, adot0, adot0 / a0, 0., 0.], dtype=torch.float64,
                                  device=dev)
        self.a, self.adot, self.hubble = self.state[0:1], self.state[1:2], self.state[2:3]
        self._sums_buf = torch.zeros(16 + 64 * 16 + 2, dtype=torch.float64, device=dev)
        self.sums = self._sums_buf[:16]
        self.energies = torch.zeros(2 * nscalars + 3, dtype=torch.float64, device=dev)

    def attach(self, stepper):
        """Advance this background inside ``stepper``'s stage kernels (their last CTA folds
        the energy sums, runs the inter-GPU exchange and the stage of ``(a, adot)``):
        nothing is launched between two stages.  :meth:`step` does this on its first call."""
        stepper.attach_expansion(self)
        return self

    def detach(self, stepper):
        if getattr(stepper, "_expansion", None) is self:
            stepper._expansion = None
        self._auto_attach = False

    def step(self, stage, stepper, dt, queue=None):
        """Advance one stage with the energy sums of the stepper's last stage (no host
        sync).  When the stage kernel has already done it in its tail (see :meth:`attach`)
        this returns immediately."""
        from pystella_b200 import _kernels
        if getattr(stepper, "_expansion", None) is self \
                and getattr(stepper, "_expanded_stage", None) == stage:
            stepper._expanded_stage = None
            sdt = stepper.dt
            if sdt is not None and abs(float(np.asarray(sdt).reshape(-1)[0]) - float(dt)) \
                    > 1e-14 * abs(float(dt)):
                raise ValueError("DeviceExpansion.step: dt differs from the stepper's dt, "
                                 "which the in-kernel background stage has already used")
            return
        if stepper.energy_partials is None or stepper.energy_nslots == 0:
            raise ValueError("the stepper has produced no energy sums: construct it with "
                             "MapKernel=FusedStencilStage, energy=True and call it first")
        nvals = stepper.energy_nvals
        if getattr(stepper, "_sums_folded", False):
            sums = stepper.energy_sums          # folded by the stage kernel's last CTA
        else:
            _kernels.energy_sums(stepper.energy_partials, stepper.energy_nslots, nvals,
                                 self._sums_buf, queue, self.strict_fp)
            sums = self.sums
        if getattr(self, "_auto_attach", True) and getattr(stepper, "fused", False) \
                and getattr(stepper, "_expansion", None) is None \
                and getattr(stepper, "_halo_mode", "local") in ("local", "peer"):
            stepper.attach_expansion(self)      # from the next stage on: in-kernel
        pg = getattr(stepper, "_peer", None)
        if pg is not None and getattr(stepper, "_halo_mode", None) == "peer":
            # all-gather over peer memory + inter-GPU barrier + expansion, one kernel
            _kernels.expansion_stage_peer(pg.args, sums, self.nscalars, self.grid_size,
                                          self.mpl, dt, self._A[stage], self._B[stage],
                                          self.state, self.energies, queue, self.strict_fp,
                                          getattr(stepper, "launch_flags", 0))
            stepper._peer_barrier_pending = False
            return
        if self.decomp is not None and self.decomp.nranks > 1:
            import torch.distributed as dist
            if sums is not self.sums:
                self.sums.copy_(sums)
                sums = self.sums
            dist.all_reduce(sums, group=self.decomp.comm)
        _kernels.expansion_stage(sums, self.nscalars, self.grid_size, self.mpl, dt,
                                 self._A[stage], self._B[stage], self.state, self.energies,
                                 queue, self.strict_fp)

    def get(self):
        """Host copy ``dict(a, adot, hubble, energies)`` (synchronises)."""
        st = self.state.cpu().numpy()
        en = self.energies.cpu().numpy()
        n = self.nscalars
        return dict(a=st[0], adot=st[1], hubble=st[2],
                    kinetic=en[:n].copy(), potential=en[n:n + 1].copy(),
                    gradient=en[n + 1:2 * n + 1].copy(), total=en[2 * n + 1],
                    pressure=en[2 * n + 2])

    def constraint(self, energy):
        """
        ``|sqrt(8 pi a^2 rho/(3 mpl^2)) a / adot - 1|`` for the energy density ``energy``
        of the CURRENT state (e.g. from :class:`Reduction`); synchronises.  Note that
        :attr:`energies` belongs to the state that *entered* the last stage, so it is
        not the right input here.
        """
        st = self.state.cpu().numpy()
        a, adot = st[0], st[1]
        return abs(np.sqrt(8 * np.pi * a**2 / 3 / self.mpl**2 * float(energy)) * a / adot - 1)
