"""
CUDA-graph capture of time steps: one launch per replay instead of one Python-driven launch
per kernel.  The reference has nothing comparable (every pyopencl launch goes through
Python, ``elementwise.py:329-353``); on a B200 small lattices are launch-bound, so this
matters as soon as a stage kernel is shorter than the Python call that issues it.

The fused stage ping-pongs the storage of ``f`` with a stepper-owned twin (one swap per
stage), so a captured sequence must contain an EVEN number of stages for the tensors to be
back in their starting buffers at the end of a replay: with an odd ``num_stages`` (e.g.
``LowStorageRK54``) capture an even number of steps.  Everything a step calls must be
stream-ordered and allocation-free after warm-up: the stage kernels, ``DeviceExpansion``,
halo kernels and the peer-memory exchange are; ``Reduction`` (host read-back) and the NCCL
halo fallback are not.
"""

import torch

__all__ = ["StepGraph", "capture_steps"]


class StepGraph:
    """Returned by :func:`capture_steps`; ``replay()`` runs the captured steps once,
    ``steps_per_replay`` says how many calls of the step function that is."""

    def __init__(self, graph, steps_per_replay, step):
        self.graph = graph
        self.steps_per_replay = steps_per_replay
        self._step = step

    def replay(self):
        self.graph.replay()

    def run(self, nsteps):
        """``nsteps`` steps: whole replays, the remainder eagerly."""
        for _ in range(nsteps // self.steps_per_replay):
            self.graph.replay()
        for _ in range(nsteps % self.steps_per_replay):
            self._step()


def capture_steps(step, steps_per_replay=2, warmup=1):
    """
    Capture ``steps_per_replay`` calls of ``step()`` (a function that issues one full time
    step on the current stream) into a CUDA graph.

    ``warmup`` eager repetitions of the same sequence run first on a side stream (lazy
    allocations, kernel attribute setup, run-time compilation), as CUDA requires.  Keep
    ``steps_per_replay`` even for steppers with an odd number of stages (see the module
    docstring).  Raises whatever the capture raises (e.g. for a step that synchronises).
    """
    if steps_per_replay < 1:
        raise ValueError("steps_per_replay must be positive")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(0, warmup)):
            for _ in range(steps_per_replay):
                step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(steps_per_replay):
            step()
    return StepGraph(graph, steps_per_replay, step)
