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

    def __init__(self, graph, steps_per_replay, step, num_stages=None):
        self.graph = graph
        self.steps_per_replay = steps_per_replay
        self.num_stages = num_stages
        self._step = step
        self._valid = True

    def replay(self):
        if not self._valid:
            raise RuntimeError(
                "this StepGraph was invalidated by an odd number of eagerly executed stages "
                "(the fused stage ping-pongs f's storage once per stage): capture again")
        self.graph.replay()

    def run(self, nsteps):
        """``nsteps`` steps: whole replays, the remainder eagerly.  An eager remainder with
        an odd number of stages leaves the live state in the other ping-pong buffer: the
        graph is then invalid and refuses further replays."""
        for _ in range(nsteps // self.steps_per_replay):
            self.replay()
        rest = nsteps % self.steps_per_replay
        for _ in range(rest):
            self._step()
        if rest and (self.num_stages is None or (rest * self.num_stages) % 2):
            self._valid = False


def capture_steps(step, steps_per_replay=2, warmup=1, num_stages=None, stepper=None):
    """
    Capture ``steps_per_replay`` calls of ``step()`` (a function that issues one full time
    step on the current stream) into a CUDA graph.

    The fused stage swaps ``f``'s storage with a twin once per stage, so a captured sequence
    must contain an EVEN number of stages; pass ``stepper=`` (or ``num_stages=``) and an odd
    ``num_stages * steps_per_replay`` is rejected (``LowStorageRK54``: capture an even number
    of steps).  Without either the caller vouches for it.

    ``warmup`` eager repetitions of the same sequence run first on a side stream (lazy
    allocations, kernel attribute setup, run-time compilation), as CUDA requires: the state
    the caller passes to ``step`` is ADVANCED by ``warmup * steps_per_replay`` steps before
    the capture (and the capture itself does not execute anything).  Raises whatever the
    capture raises (e.g. for a step that synchronises).
    """
    if steps_per_replay < 1:
        raise ValueError("steps_per_replay must be positive")
    if stepper is not None and num_stages is None:
        num_stages = stepper.num_stages
    if num_stages is not None and (num_stages * steps_per_replay) % 2:
        raise ValueError(
            f"{steps_per_replay} step(s) of {num_stages} stages is an odd number of stages: "
            "the ping-pong storage of the fused stage would not be back in place at the end "
            "of a replay; capture an even number of steps")
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
    return StepGraph(graph, steps_per_replay, step, num_stages)
