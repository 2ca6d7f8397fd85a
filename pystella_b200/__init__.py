"""
pystella_b200 -- a B200-native lattice-PDE stepper with pystella's API surface.

The user-facing classes mirror zachjweiner/pystella (``pystella/__init__.py:24-63``)
for the time-stepping hot path; device arrays are ``torch.Tensor`` on CUDA and a
``queue`` is a ``torch.cuda.Stream`` (or *None* for the current stream).  Every
device operation goes through the C ABI of ``libpystella_b200.so``
(``include/pystella_b200.h``); there is no CPU fallback.
"""

from pystella_b200.field import (Field, DynamicField, index_fields, shift_fields,
                                 substitute, diff, get_field_args, var)
from pystella_b200.sectors import (Sector, ScalarSector, TensorPerturbationSector,
                                   tensor_index, get_rho_and_p)
from pystella_b200.step import (
    ElementWiseMap, FusedStencilStage, Stencil, StreamingStencil, Stepper,
    RungeKutta4, RungeKutta3SSP, RungeKutta3Heun, RungeKutta3Nystrom,
    RungeKutta3Ralston, RungeKutta2Midpoint, RungeKutta2Heun, RungeKutta2Ralston,
    LowStorageRK54, LowStorageRK144, LowStorageRK134, LowStorageRK124,
    LowStorageRK3Williamson, LowStorageRK3Inhomogeneous, LowStorageRK3Symmetric,
    LowStorageRK3PredictorCorrector, LowStorageRK3SSP, all_steppers)
from pystella_b200.derivs import (FiniteDifferencer, FirstCenteredDifference,
                                  SecondCenteredDifference, expand_stencil,
                                  centered_diff)
from pystella_b200.decomp import DomainDecomposition
from pystella_b200.reduction import Reduction, FieldStatistics
from pystella_b200.expansion import Expansion, DeviceExpansion
from pystella_b200.fourier import DFT, RayleighGenerator, PowerSpectra
from pystella_b200.projectors import Projector
from pystella_b200.spectral import SpectralCollocator, SpectralPoissonSolver
from pystella_b200.histogram import Histogrammer, FieldHistogrammer
from pystella_b200 import output
from pystella_b200.graph import StepGraph, capture_steps
from pystella_b200.output import OutputFile
from pystella_b200._lib import PsbError

__version__ = "0.1.0"


def choose_device(local_rank=None):
    """
    Bind this process to its GPU (one process per GPU): the analogue of
    ``pystella.choose_device_and_make_context`` (``pystella/__init__.py:66-102``).
    Uses ``LOCAL_RANK`` (torchrun) unless ``local_rank`` is given.
    """
    import os
    import torch
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    return torch.device("cuda", local_rank)


def choose_device_and_make_context(platform_choice=None, device_choice=None):
    """The reference's entry point (``pystella/__init__.py:45-102``): binds the process to
    its GPU and returns the ``torch.device`` (there is no context object to make)."""
    return choose_device(device_choice if isinstance(device_choice, int) else None)


__all__ = [
    "Field", "DynamicField", "index_fields", "shift_fields", "substitute", "diff",
    "get_field_args", "var",
    "DFT", "RayleighGenerator", "PowerSpectra", "Projector",
    "SpectralCollocator", "SpectralPoissonSolver", "Histogrammer", "FieldHistogrammer",
    "OutputFile", "StepGraph", "capture_steps",
    "Sector", "ScalarSector", "TensorPerturbationSector", "tensor_index",
    "get_rho_and_p",
    "ElementWiseMap", "FusedStencilStage", "Stencil", "StreamingStencil", "Stepper",
    "choose_device_and_make_context",
    "RungeKutta4", "RungeKutta3SSP", "RungeKutta3Heun", "RungeKutta3Nystrom",
    "RungeKutta3Ralston", "RungeKutta2Midpoint", "RungeKutta2Heun",
    "RungeKutta2Ralston",
    "LowStorageRK54", "LowStorageRK144", "LowStorageRK134", "LowStorageRK124",
    "LowStorageRK3Williamson", "LowStorageRK3Inhomogeneous", "LowStorageRK3Symmetric",
    "LowStorageRK3PredictorCorrector", "LowStorageRK3SSP", "all_steppers",
    "FiniteDifferencer", "FirstCenteredDifference", "SecondCenteredDifference",
    "expand_stencil", "centered_diff",
    "DomainDecomposition", "Reduction", "FieldStatistics", "Expansion",
    "DeviceExpansion",
    "PsbError", "choose_device",
]
