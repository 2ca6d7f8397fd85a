"""
Worker run by the multi-process tests (gloo on CPU, nccl on GPUs):
    python tests/dist_worker.py <case> <world> <port> [args...]
Spawns `world` ranks with torch.multiprocessing and exits non-zero on any failure.
"""
import os
import sys
import traceback

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _decomp_case(rank, world, proc_shape, h, grid_shape, backend):
    import pystella_b200 as ps
    from helpers import HostOps, periodic_global
    import oracle as O

    hh = (h,) * 3 if isinstance(h, int) else tuple(h)
    decomp = ps.DomainDecomposition(proc_shape, h, grid_shape=grid_shape)
    cuda = backend == "nccl"
    if not cuda:
        decomp.host_ops = HostOps()
    assert decomp.nranks == world and decomp.rank == rank
    assert decomp.rank_tuple == O.rank_tuple_of(proc_shape, rank)
    rank_shape, start = decomp.get_rank_shape_start(grid_shape)
    assert (rank_shape, start) == O.rank_shape_start(proc_shape, grid_shape,
                                                     decomp.rank_tuple)
    assert decomp.rank_shape == rank_shape

    # same data on every rank (test/test_decomp.py:61-76)
    rng = np.random.default_rng(12321)
    data = periodic_global(rng, grid_shape, hh)
    sub = np.zeros(tuple(n + 2 * x for n, x in zip(rank_shape, hh)))
    rank_slice = tuple(slice(s + x, s + n + x) for n, s, x in zip(rank_shape, start, hh))
    unpadded = tuple(slice(x, -x) if x > 0 else slice(None) for x in hh)
    sub[unpadded] = data[rank_slice]
    t = torch.from_numpy(sub.copy())
    if cuda:
        t = t.cuda()
    decomp.share_halos(None, t)
    if cuda:
        torch.cuda.synchronize()
    pencil = tuple(slice(s, s + n + 2 * x) for n, s, x in zip(rank_shape, start, hh))
    got = t.cpu().numpy()
    assert np.array_equal(got, data[pencil]), f"rank {rank} {decomp.rank_tuple}: halos wrong"

    # collectives
    assert decomp.allreduce(rank + 1) == world * (world + 1) // 2
    assert decomp.allreduce(float(rank), op="max") == float(world - 1)
    assert decomp.bcast("hello" if rank == 1 else None, 1) == "hello"
    decomp.Barrier()

    # gather / scatter round trip (test/test_decomp.py:108-173)
    full = np.zeros(grid_shape) if rank == 0 else None
    decomp.gather_array(None, t if cuda else t.numpy(), full, 0)
    if rank == 0:
        interior = tuple(slice(x, x + n) for x, n in zip(hh, grid_shape))
        assert np.array_equal(full, data[interior])
    back = np.zeros_like(sub)
    decomp.scatter_array(None, full, back, 0)
    assert np.array_equal(back[unpadded], sub[unpadded])


def _reduction_case(rank, world, backend):
    """Reduction / FieldStatistics over ranks (nccl only: kernels are CUDA)."""
    import pystella_b200 as ps
    grid = (32, 16, 24)
    h = 2
    decomp = ps.DomainDecomposition((world, 1, 1), h, grid_shape=grid)
    rs, start = decomp.get_rank_shape_start(grid)
    rng = np.random.default_rng(7)
    full = rng.standard_normal(grid)
    sub = np.zeros(tuple(n + 2 * h for n in rs))
    sub[h:-h, h:-h, h:-h] = full[start[0]:start[0] + rs[0]]
    t = torch.from_numpy(sub).cuda()
    stats = ps.FieldStatistics(decomp, h, max_min=True)
    out = stats(t)
    assert np.isclose(out["mean"], full.mean(), rtol=1e-12)
    assert np.isclose(out["variance"], full.var(), rtol=1e-11)
    assert out["max"] == full.max() and out["min"] == full.min()
    red = ps.Reduction(decomp, {"s": [(ps.Field("f", offset="h"), "sum")]}, halo_shape=h)
    assert np.isclose(red(None, f=t)["s"][0], full.sum(), rtol=1e-12, atol=1e-12)


def _fused_case(rank, world, backend, nscalars, energy, mode="peer"):
    """Multi-rank fused stepper (boundary/interior split + face exchange) == single rank."""
    peer_tail = mode == "peertail"     # in-kernel inter-GPU exchange (watcher CTA)
    if peer_tail:
        mode = "peer"
    os.environ["PSB_HALO_MODE"] = mode
    import pystella_b200 as ps
    import oracle as O
    h = 2
    # 40 planes per rank: boundary chunks + interior in NCCL mode; uneven in peer mode
    grid = (40 * world + (3 if mode == "peer" else 0), 16, 32)
    dx = tuple(5 / n for n in grid)
    dt = 0.1 * min(dx)
    rng = np.random.default_rng(99)
    pad_full = tuple(n + 2 * h for n in grid)
    f0 = 0.2 + 1e-2 * rng.standard_normal((nscalars,) + pad_full)
    d0 = -0.1 + 1e-2 * rng.standard_normal((nscalars,) + pad_full)
    if nscalars == 1:
        terms = [(0.5, (2,))]
        def potential(f):
            return f[0]**2 / 2
    else:
        terms = [(0.5, (2, 0)), (0.5, (0, 2)), (0.5, (2, 2))]
        def potential(f):
            return f[0]**2 / 2 + f[1]**2 / 2 + f[0]**2 * f[1]**2 / 2

    # global oracle
    fr, dr = f0.copy(), d0.copy()
    kf, kd = np.zeros_like(fr), np.zeros_like(fr)
    lap = np.zeros((nscalars,) + grid)
    nsteps = 2
    for _ in range(nsteps):
        for s in range(5):
            O.reference_schedule_stage(s, fr, dr, kf, kd, lap, h, dx, dt, a=1.05, hubble=0.2,
                                       potential_terms=terms)

    decomp = ps.DomainDecomposition((world, 1, 1), h, grid_shape=grid)
    rs, start = decomp.get_rank_shape_start(grid)
    sl = slice(start[0], start[0] + rs[0] + 2 * h)
    f = torch.from_numpy(np.ascontiguousarray(f0[:, sl])).cuda()
    d = torch.from_numpy(np.ascontiguousarray(d0[:, sl])).cuda()
    sector = ps.ScalarSector(nscalars, potential=potential)
    st = ps.LowStorageRK54(sector, halo_shape=h, dt=dt, dx=dx, decomp=decomp,
                           MapKernel=ps.FusedStencilStage, energy=energy,
                           tune={"peer_tail": peer_tail})
    dexp = None
    for _ in range(nsteps):
        for s in range(5):
            st(s, f=f, dfdt=d, a=1.05, hubble=0.2)
    torch.cuda.synchronize()
    assert st._halo_mode == mode, (st._halo_mode, mode)
    inner = slice(start[0] + h, start[0] + h + rs[0])
    for n in range(nscalars):
        g = f[n].cpu().numpy()[h:-h, h:-h, h:-h]
        r = fr[n][inner, h:-h, h:-h]
        err = np.max(np.abs(g - r)) / np.max(np.abs(r))
        assert err < 1e-12, f"rank {rank} field {n}: {err}"
        g = d[n].cpu().numpy()[h:-h, h:-h, h:-h]
        r = dr[n][inner, h:-h, h:-h]
        assert np.max(np.abs(g - r)) / np.max(np.abs(r)) < 1e-12
    if energy:
        # energy sums of the last stage's input state, all-reduced, vs. a global reduction
        # energies of the last stage's INPUT state: reproduce with the oracle by
        # re-running everything but that stage
        fr2, dr2 = f0.copy(), d0.copy()
        kf2, kd2 = np.zeros_like(fr2), np.zeros_like(fr2)
        for i in range(nsteps * 5 - 1):
            O.reference_schedule_stage(i % 5, fr2, dr2, kf2, kd2, lap, h, dx, dt, a=1.05,
                                       hubble=0.2, potential_terms=terms)
        fi, di, laps = [], [], []
        for n in range(nscalars):
            g = fr2[n].copy()
            O.share_halos(g, h)
            laps.append(O.laplacian(g, h, dx))
            fi.append(O.interior(fr2[n], h))
            di.append(O.interior(dr2[n], h))
        ref = O.get_rho_and_p(O.scalar_energy(fi, di, laps, 1.3, terms))
        dexp = ps.DeviceExpansion(1.0, ps.LowStorageRK54, grid_size=float(np.prod(grid)),
                                  decomp=decomp, nscalars=nscalars, a0=1.3)
        dexp.step(0, st, dt)
        got = dexp.get()
        assert np.isclose(got["total"], ref["total"], rtol=1e-12), (got["total"], ref["total"])
        assert np.isclose(got["pressure"], ref["pressure"], rtol=1e-10, atol=1e-14)
        # identical on all ranks, bit for bit
        tot = torch.tensor([got["total"]], dtype=torch.float64, device="cuda")
        lo, hi = tot.clone(), tot.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float(lo) == float(hi)


def _fused_gw_case(rank, world, backend, dtype, peer_tail=False):
    """Config-4 structure on several ranks: fused scalar stage + fused tensor stage, both with
    peer-memory halos, the stage's inter-GPU barrier and the background in the tensor kernel's
    watcher CTA == the C oracle on the global lattice."""
    import pystella_b200 as ps
    import oracle as O
    from oracle import c_oracle as CO
    h = 2
    grid = (20 * world + 4, 12, 16)          # uneven slabs
    dx = tuple(5 / n for n in grid)
    dt = 0.1 * min(dx)
    rng = np.random.default_rng(7)
    pad = tuple(n + 2 * h for n in grid)
    npdt = np.float64 if dtype == "f64" else np.float32
    f0 = (0.2 + 1e-2 * rng.standard_normal((2,) + pad)).astype(npdt)
    d0 = (-0.1 + 1e-2 * rng.standard_normal((2,) + pad)).astype(npdt)
    h0 = (1e-3 * rng.standard_normal((6,) + pad)).astype(npdt)
    dh0 = (1e-3 * rng.standard_normal((6,) + pad)).astype(npdt)
    terms = [(0.5, (2, 0)), (0.3, (2, 2))]
    gs = float(np.prod(grid))

    # ---- oracle: global lattice, reference loop with the background
    F, D, Hh, DH = (x.astype(np.float64) for x in (f0, d0, h0, dh0))
    kF, kD, kH, kDH = (np.zeros_like(x) for x in (F, D, Hh, DH))
    pot = CO.make_potential(2, terms)
    lap_f, dfdx, lap_hij = CO.gw_initial_derivatives(F, Hh, h, dx)
    en = CO.scalar_energy(F, D, lap_f, h, 1.0, gs, pot)
    e0 = 0.37
    adot = float(np.sqrt(8 * np.pi / 3 * e0))
    state = np.array([1.0, adot, adot, 0.0, 0.0])
    A, B, _ = O.LSRK_TABLES["LowStorageRK54"]
    nsteps = 2
    CO.preheating_gw_steps(nsteps, A, B, F, D, kF, kD, lap_f, dfdx, Hh, DH, kH, kDH, lap_hij, h,
                           dx, dt, 1.0, gs, pot, state, en, True)

    # ---- device, x-slabs
    decomp = ps.DomainDecomposition((world, 1, 1), h, grid_shape=grid)
    rs, start = decomp.get_rank_shape_start(grid)
    sl = slice(start[0], start[0] + rs[0] + 2 * h)
    arrays = dict(f=f0, dfdt=d0, hij=h0, dhijdt=dh0)
    arrays = {k: torch.from_numpy(np.ascontiguousarray(v[:, sl])).cuda()
              for k, v in arrays.items()}
    scalar = ps.ScalarSector(2, potential=lambda g: g[0]**2 / 2 + 0.3 * g[0]**2 * g[1]**2)
    gw = ps.TensorPerturbationSector([scalar])
    st = ps.LowStorageRK54([scalar, gw], halo_shape=h, dt=dt, dx=dx, decomp=decomp,
                           MapKernel=ps.FusedStencilStage, energy=True,
                           tune={"peer_tail": peer_tail})
    dexp = ps.DeviceExpansion(e0, ps.LowStorageRK54, grid_size=gs, decomp=decomp, nscalars=2)
    dexp.attach(st)
    for _ in range(nsteps):
        for s in range(5):
            st(s, a=dexp.a, hubble=dexp.hubble, **arrays)
            dexp.step(s, st, dt)
    torch.cuda.synchronize()
    assert st._halo_mode == "peer"
    got = dexp.get()
    tol = 1e-12 if dtype == "f64" else 3e-5
    assert np.isclose(got["a"], state[0], rtol=1e-13 if dtype == "f64" else 1e-6)
    assert np.isclose(got["hubble"], state[2], rtol=1e-11 if dtype == "f64" else 1e-5)
    inner = slice(start[0] + h, start[0] + h + rs[0])
    for name, ref in (("f", F), ("dfdt", D), ("hij", Hh), ("dhijdt", DH)):
        for c in range(ref.shape[0]):
            g = arrays[name][c].cpu().numpy()[h:-h, h:-h, h:-h].astype(np.float64)
            r = ref[c][inner, h:-h, h:-h]
            err = np.max(np.abs(g - r)) / np.max(np.abs(ref[c]))
            assert err < tol, f"rank {rank} {name}[{c}]: {err}"


def _fourier_case(rank, world, backend, proc_shape=None):
    """Slab-distributed DFT / RayleighGenerator / PowerSpectra == the single-rank results
    on the gathered arrays (test/test_dft.py on more than one rank)."""
    import pystella_b200 as ps
    from oracle import fourier_oracle as F
    cuda = backend == "nccl"
    fft_backend = None if cuda else "fftw"
    for grid, dtype in (((16, 12, 10), "float64"), ((9, 7, 8), "float64"),
                        ((8, 6, 4), "complex128")):
        h = 1
        decomp = ps.DomainDecomposition(proc_shape or (world, 1, 1), h, grid_shape=grid)
        fft = ps.DFT(decomp, None, None, grid, dtype, backend=fft_backend)
        rs, start = decomp.get_rank_shape_start(grid)
        assert fft.shape(False) == rs
        rng = np.random.default_rng(5)
        full = rng.random(grid)
        if not fft.is_real:
            full = full + 1j * rng.random(grid)
        loc = full[start[0]:start[0] + rs[0], start[1]:start[1] + rs[1]].copy()
        ref = np.fft.rfftn(full) if fft.is_real else np.fft.fftn(full)
        ys, ny = fft._y_start, fft.shape(True)[1]
        zs, nzk = fft._zk_start, fft.shape(True)[2]
        ref = ref[:, :, zs:zs + nzk]
        assert sum(fft._y_sizes) == grid[1]
        arg = torch.from_numpy(loc).cuda() if cuda else loc
        fk = fft.dft(arg)
        assert tuple(fk.shape) == fft.shape(True)
        got = fk.cpu().numpy().copy()
        assert np.allclose(got, ref[:, ys:ys + ny], rtol=1e-11, atol=1e-11), (rank, grid)
        # momenta are the local slices
        ky = F.fftfreq(grid[1])[ys:ys + ny]
        assert np.array_equal(fft.sub_k["momenta_y"].cpu().numpy(), ky)
        # round trip, padded output
        pad = np.zeros(tuple(n + 2 * h for n in rs), dtype)
        out = torch.from_numpy(pad).cuda() if cuda else pad
        fft.idft(fk, out)
        back = out.cpu().numpy() if cuda else out
        assert np.allclose(back[h:-h, h:-h, h:-h], loc * np.prod(grid), rtol=1e-11,
                           atol=1e-11)
        if not fft.is_real:
            continue
        # spectra: same bins and values as a single-rank evaluation of the full array
        L = (5., 6., 7.)
        dk = tuple(2 * np.pi / Li for Li in L)
        spec = ps.PowerSpectra(decomp, fft, dk, float(np.prod(L)))
        one = _single_rank_spectrum(full, grid, dk, float(np.prod(L)))
        assert spec.num_bins == one[0] and np.array_equal(spec.bin_counts, one[1])
        mine = spec(arg, k_power=2)
        ok = one[1] > 0
        assert np.allclose(mine[ok], one[2][ok], rtol=1e-11), rank
        # Rayleigh modes: local slices, real field after the backward transform
        modes = ps.RayleighGenerator(None, fft, dk, float(np.prod(L)), seed=11 * (rank + 1))
        gk, dgk = modes.generate_WKB(None, hubble=0.1,
                                     omega_k=lambda k: np.sqrt(k**2 + 0.5))
        assert tuple(gk.shape) == fft.shape(True)
        if ys == 0 and zs == 0:
            assert complex(gk[0, 0, 0]) == 0        # homogeneous mode carries no power
        fx = fft.idft(gk).cpu().numpy()
        assert fx.shape == rs and np.all(np.isfinite(fx)) and np.std(fx) > 0


def _single_rank_spectrum(full, grid, dk, volume):
    """PowerSpectra of the gathered array, evaluated without any decomposition."""
    from oracle import fourier_oracle as F
    fk = np.fft.rfftn(full)
    kvecs = np.meshgrid(*F.momenta(grid), indexing="ij")
    kmags = np.sqrt(sum((d * k)**2 for d, k in zip(dk, kvecs)))
    counts = 2. * np.ones_like(kmags)
    counts[kvecs[2] == 0] = 1.
    counts[kvecs[2] == grid[-1] // 2] = 1.
    bw = min(dk)
    num_bins = int(np.max(kmags) / bw + .5) + 1
    edges = np.arange(-.5, num_bins + .5) * bw
    bin_counts = np.histogram(kmags, weights=counts, bins=edges)[0]
    mult = np.where((kvecs[2] > 0) & (kvecs[2] < grid[-1] / 2), 2., 1.)
    hist = np.histogram(kmags, weights=mult * kmags**2 * np.abs(fk)**2, bins=edges)[0]
    norm = (1 / 2 / np.pi**2 / volume) * (volume / np.prod(grid))**2
    with np.errstate(invalid="ignore", divide="ignore"):
        return num_bins, bin_counts, norm * hist / bin_counts


def _install_shared_memory_peers(mp_):
    """Peer memory between the ranks of one node, for the emulated device: POSIX shared
    memory takes the place of NVLink-mapped allocations.  `torch.distributed.
    _symmetric_memory` (which pystella_b200/peer.py uses for allocation and address exchange
    only) is replaced by a module with the same three calls: `empty` allocates a named
    segment, `rendezvous` exchanges the names over the process group and maps the peers'
    segments, `get_buffer(q, ...)` is this process's view of rank q's segment.  The kernels
    of the ranks run concurrently in their own processes, so the flag barriers of
    psb_expansion_stage_peer / the in-kernel tail really wait on one another."""
    import atexit
    import types
    from multiprocessing import shared_memory

    owned, mapped = {}, []

    def empty(*size, dtype=torch.float64, device=None):
        numel = int(np.prod(size)) if size else 1
        nbytes = max(1, numel * torch.empty(0, dtype=dtype).element_size())
        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        t = torch.frombuffer(shm.buf, dtype=dtype, count=numel)
        t.zero_()
        owned[t.data_ptr()] = (shm, numel, dtype)
        return t

    class Handle:
        def __init__(self, views):
            self.views = views

        def barrier(self, *a, **k):
            dist.barrier()

        def get_buffer(self, q, shape, dtype, *a, **k):
            n = int(np.prod(shape))
            return self.views[q].view(dtype)[:n].view(tuple(shape))

    def rendezvous(buf, group_name=None):
        shm, numel, dtype = owned[buf.data_ptr()]
        names = [None] * dist.get_world_size()
        dist.all_gather_object(names, (shm.name, numel))
        views = []
        for q, (name, n) in enumerate(names):
            if q == dist.get_rank():
                views.append(buf)
                continue
            peer = shared_memory.SharedMemory(name=name)
            mapped.append(peer)
            views.append(torch.frombuffer(peer.buf, dtype=dtype, count=n))
        dist.barrier()
        return Handle(views)

    def cleanup():
        for shm, _, _ in owned.values():
            try:
                shm.unlink()
            except Exception:  # noqa: BLE001
                pass
    atexit.register(cleanup)

    fake = types.ModuleType("torch.distributed._symmetric_memory")
    fake.empty, fake.rendezvous = empty, rendezvous
    sys.modules["torch.distributed._symmetric_memory"] = fake
    torch.distributed._symmetric_memory = fake
    from pystella_b200 import peer
    mp_.setattr(peer, "peer_memory_available",
                lambda decomp: decomp is not None and decomp.nranks > 1)


def _emulate_cuda():
    """Case `emu_fused`: the multi-rank fused stepper on gloo ranks WITHOUT a GPU -- host
    tensors, the CUDA kernels executed on the CPU model (tests/emu_device.py), streams and
    events reduced to program order.  -> the object that keeps the patches alive."""
    import contextlib
    import types
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from emu_device import emulated_device

    class Patches:
        def __init__(self):
            self.undo = []

        def setattr(self, obj, name, val):
            self.undo.append((obj, name, getattr(obj, name)))
            setattr(obj, name, val)

    mp_ = Patches()
    stack = contextlib.ExitStack()
    dev = stack.enter_context(emulated_device(mp_))
    dev.main.emu_set_reverse_blocks(1)
    os.environ["PSB_JIT"] = "0"          # coefficient tables (generated instances: GPU only)

    class Stream:
        def __init__(self, *a, **k):
            pass

        def wait_event(self, ev):
            pass

        def wait_stream(self, st):
            pass

    class Event:
        def __init__(self, *a, **k):
            pass

        def record(self, stream=None):
            pass

        def synchronize(self):
            pass

    def strip(fn):
        def wrapped(*a, **k):
            k.pop("device", None)
            return fn(*a, **k)
        return wrapped

    mp_.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self.clone())   # a copy, like H2D
    _install_shared_memory_peers(mp_)
    for name in ("zeros", "empty", "ones", "tensor", "full"):
        mp_.setattr(torch, name, strip(getattr(torch, name)))
    mp_.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    mp_.setattr(torch.cuda, "current_device", lambda: 0)
    mp_.setattr(torch.cuda, "Stream", Stream)
    mp_.setattr(torch.cuda, "Event", Event)
    mp_.setattr(torch.cuda, "current_stream", lambda *a, **k: Stream())
    mp_.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    return types.SimpleNamespace(stack=stack, patches=mp_, dev=dev)


def _worker(rank, world, port, case, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    backend = "nccl" if case.startswith("gpu") else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group(backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        if case in ("cpu_decomp", "gpu_decomp"):
            proc_shape = tuple(int(x) for x in args[0].split(","))
            for h, grid in ((1, (32, 32, 32)), (2, (19, 23, 16)), ((2, 0, 3), (32, 32, 32)),
                            ((0, 2, 1), (32, 32, 32)), (1, (109, 67, 32))):
                _decomp_case(rank, world, proc_shape, h, grid, backend)
        elif case in ("cpu_fourier", "gpu_fourier"):
            proc_shape = tuple(int(x) for x in args[0].split(",")) if args else None
            _fourier_case(rank, world, backend, proc_shape)
        elif case == "gpu_reduction":
            _reduction_case(rank, world, backend)
        elif case == "gpu_fused_gw":
            _fused_gw_case(rank, world, backend, args[0], len(args) > 1 and args[1] == "1")
        elif case == "gpu_fused":
            _fused_case(rank, world, backend, int(args[0]), args[1] == "1",
                        args[2] if len(args) > 2 else "peer")
        elif case == "emu_fused_gw":
            keep = _emulate_cuda()      # noqa: F841
            _fused_gw_case(rank, world, backend, args[0], len(args) > 1 and args[1] == "1")
        elif case == "emu_fused":
            keep = _emulate_cuda()      # noqa: F841 -- patches live as long as the worker
            _fused_case(rank, world, backend, int(args[0]), args[1] == "1",
                        args[2] if len(args) > 2 else "nccl")
        else:
            raise ValueError(case)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def main():
    case, world, port = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    args = sys.argv[4:]
    try:
        mp.spawn(_worker, args=(world, port, case, args), nprocs=world, join=True)
    except Exception:
        traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
