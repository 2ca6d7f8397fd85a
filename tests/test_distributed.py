"""
Multi-rank paths.  CPU (gloo, world_size 2-4): rank arithmetic, message ordering of
share_halos (incl. (Px, Py, 1) decompositions with packed y faces and uneven
splits), collectives, gather/scatter -- the halo kernels are replaced by the
oracle-backed tests/helpers.HostOps there.  GPU (nccl): the same through the CUDA
kernels, reductions over ranks, and the fused stepper's boundary/interior schedule
against a single global oracle run.
"""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run_group(cmd, timeout):
    """Run `cmd` in its own process group; on timeout the whole group (spawned ranks
    included) is killed, so that nothing keeps spinning on a GPU afterwards."""
    import signal
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                            start_new_session=True)
    try:
        out, err = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        return 124, out, err + f"\n[timed out after {timeout} s; process group killed]"
    return proc.returncode, out, err


def _run(case, world, *args, timeout=300):
    for attempt in range(3):
        cmd = [sys.executable, os.path.join(HERE, "dist_worker.py"), case, str(world),
               str(_free_port()), *args]
        rc, out, err = run_group(cmd, timeout)
        if rc == 0:
            return
        if "EADDRINUSE" not in err + out:
            break   # a real failure; port collisions are retried
    assert rc == 0, out[-3000:] + err[-3000:]


@pytest.mark.parametrize("proc_shape", ["2,1,1", "1,2,1", "2,2,1", "3,1,1"])
def test_decomposition_gloo(proc_shape):
    world = 1
    for p in proc_shape.split(","):
        world *= int(p)
    _run("cpu_decomp", world, proc_shape)


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_fourier_gloo(world):
    """Slab DFT (all-to-all transposes), spectra and Rayleigh modes over gloo."""
    _run("cpu_fourier", world)


@pytest.mark.parametrize("proc_shape", ["2,2,1", "1,2,1"])
def test_pencil_fourier_gloo(proc_shape):
    """Pencil DFT: transposes inside the row and the column communicator."""
    world = 1
    for p in proc_shape.split(","):
        world *= int(p)
    _run("cpu_fourier", world, proc_shape)


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("proc_shape", ["2,1,1", "1,2,1"])
def test_decomposition_nccl(proc_shape):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    _run("gpu_decomp", 2, proc_shape)


@pytest.mark.gpu
def test_reduction_nccl():
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    _run("gpu_reduction", 2)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["peer", "peertail", "nccl"])
@pytest.mark.parametrize("nscalars,energy", [(1, "0"), (2, "1")])
def test_fused_stepper_two_ranks(nscalars, energy, mode):
    """peer: halo planes stored into the neighbour's memory by the stage kernel +
    peer all-gather/barrier kernel; peertail: the same exchange inside the stage kernel's
    watcher CTA; nccl: boundary-first schedule with send/recv."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    _run("gpu_fused", 2, str(nscalars), energy, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_fused_stepper_four_ranks(mode):
    if _ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    _run("gpu_fused", 4, "1", "1", mode)


@pytest.mark.gpu
@pytest.mark.parametrize("peer_tail", ["0", "1"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_fused_scalar_plus_tensor_two_ranks(dtype, peer_tail):
    """Config-4 structure over peer memory: scalar + tensor stage kernels store their
    boundary planes into the neighbour's halos, the tensor kernel's watcher CTA runs the
    inter-GPU barrier and the background stage; == the C oracle on the global lattice."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    _run("gpu_fused_gw", 2, dtype, peer_tail)
