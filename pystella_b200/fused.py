"""
Orchestration of the fused stencil + RK stage (``psb_scalar_stage_fused``).

Per stage and scalar system:

* single rank: one launch over the whole slab; the kernel's store epilogue also
  writes the periodic images onto the face halos of the new ``f``;
* ``proc_shape = (P, 1, 1)``, P > 1: the two boundary slabs (the ``h`` outermost
  x-planes on either side) are updated first; their new faces are then sent to
  the x-neighbours on a side stream (NCCL over NVLink, straight out of / into the
  field tensor -- x faces are contiguous) while the interior slab is updated on
  the compute stream; the streams join before the next stage.  This replaces the
  reference's synchronous pack -> D2H -> MPI.Sendrecv -> H2D -> unpack
  (``pystella/decomp.py:402-419``) called from ``FiniteDifferencer.__call__``
  (``pystella/derivs.py:412``).

The new ``f`` is written to a stepper-owned twin buffer (neighbouring sites still
need the old values); afterwards the *storage* of the user's tensor and of the twin
are swapped (no copy), so the user-visible contract "``f`` is updated in place"
(``examples/wave_equation.py:64``) holds for the tensor object passed in.  Views
of ``f`` taken before the call keep pointing at the old storage.
"""

import os

import numpy as np
import torch
import torch.distributed as dist

from pystella_b200 import _kernels

__all__ = ["fused_scalar_stage"]


def _swap_storage(a, b):
    a_data, b_data = a.data, b.data
    a.data, b.data = b_data, a_data


def halos_are_fresh(stepper, key, f):
    """Were the halos of ``f`` last written by this stepper's own stage kernel?  Tracked per
    system: the tensor object, its storage and its version counter must all be the ones the
    stepper left behind (a different tensor, a re-uploaded field or an in-place edit between
    steps -- rescaling, checkpoint load -- all get a fresh ``share_halos``, as the reference
    does on every derivative call, ``pystella/derivs.py:412``)."""
    rec = stepper.__dict__.setdefault("_fresh", {}).get(key)
    return (rec is not None and not stepper.__dict__.get("_force_refresh", False)
            and rec[0]() is f and rec[1] == f.data_ptr() and rec[2] == f._version)


def mark_halos_fresh(stepper, key, f):
    import weakref
    stepper.__dict__.setdefault("_fresh", {})[key] = (weakref.ref(f), f.data_ptr(),
                                                      f._version)


def fill_halos(stepper, f, h, queue):
    """share_halos of every field of padded ``f`` (rank-local wrap or exchange)."""
    decomp = stepper.decomp
    views = [f[n] for n in range(f.shape[0])] if f.dim() == 4 else [f]
    for v in views:
        if decomp is not None:
            decomp.share_halos(queue, v)
        else:
            _kernels.halo_wrap(v, (h, h, h), 7, queue)


TIMELINE = os.environ.get("PSB_TIMELINE", "0") == "1"
_timeline = []   # (label, event) tuples appended in issue order when PSB_TIMELINE=1


def _mark(label, stream):
    if TIMELINE:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream)
        _timeline.append((label, ev))


def timeline_report():
    """Average gaps between consecutive marks (ms) since the last report."""
    torch.cuda.synchronize()
    out = {}
    base = None
    for label, ev in _timeline:
        if label == "stage_begin":
            base = ev
            continue
        if base is not None:
            out.setdefault(label, []).append(base.elapsed_time(ev))
    _timeline.clear()
    return {k: sum(v) / len(v) for k, v in out.items()}


def _ensure_streams(stepper):
    if stepper._streams is None:
        stepper._streams = {
            "comm": torch.cuda.Stream(priority=-1),
            "boundary_done": torch.cuda.Event(),
            "comm_done": torch.cuda.Event(),
        }
    return stepper._streams


def exchange_x_faces(decomp, f, h, fields_axis):
    """
    Post the x-face exchange of padded ``f`` (shape ``([n,] Px, Py, Pz)``): my low
    interior planes go to rank ``rx-1``'s high halo and my high interior planes to
    rank ``rx+1``'s low halo (``pystella/decomp.py:405-415``).  One batched group
    of sends/receives for all fields.
    """
    lo = decomp.rankID(decomp.rx - 1, decomp.ry, decomp.rz)
    hi = decomp.rankID(decomp.rx + 1, decomp.ry, decomp.rz)
    ops = []
    views = [f[n] for n in range(f.shape[0])] if fields_axis else [f]
    for v in views:
        Nx = v.shape[0] - 2 * h
        ops += [
            dist.P2POp(dist.isend, v[h:2 * h], lo, group=decomp.comm),
            dist.P2POp(dist.irecv, v[Nx + h:], hi, group=decomp.comm),
            dist.P2POp(dist.isend, v[Nx:Nx + h], hi, group=decomp.comm),
            dist.P2POp(dist.irecv, v[:h], lo, group=decomp.comm),
        ]
    return dist.batch_isend_irecv(ops)


def _launch_ranges(Nx, h, multi, chunk=16):
    """x-ranges of the launches of one stage: whole slab, or lo / hi boundary slabs
    followed by the interior."""
    if not multi:
        return [(0, Nx)]
    # the boundary launch takes whole x-chunks (no extra prologue planes compared
    # with being part of the interior launch); the faces are their outermost h planes
    hb = min(max(h, chunk), Nx)
    if Nx - hb > hb:
        # both boundary slabs in ONE launch (two x ranges), then the interior
        return [((0, hb), (Nx - hb, Nx)), (hb, Nx - hb)]
    return [(0, Nx)]   # slab so thin that everything is boundary


def _halo_mode(stepper, decomp, multi):
    """'peer' (stores over NVLink from inside the stage kernel) or 'nccl' (send/recv
    on a side stream overlapped with the interior launch)."""
    if not multi:
        return "local"
    mode = getattr(stepper, "_halo_mode", None)
    if mode is None:
        from pystella_b200.peer import PeerGroup, peer_memory_available
        want = os.environ.get("PSB_HALO_MODE", "peer")
        mode = "nccl"
        if want == "peer" and peer_memory_available(decomp):
            try:
                stepper._peer = PeerGroup(decomp, torch.device("cuda",
                                                               torch.cuda.current_device()))
                mode = "peer"
            except Exception as exc:  # noqa: BLE001
                import logging
                logging.getLogger(__name__).warning(
                    "peer memory unavailable (%r): falling back to NCCL halos", exc)
        # all ranks must agree
        agreed = decomp.allreduce(1 if mode == "peer" else 0, op="min")
        mode = "peer" if agreed == 1 else "nccl"
        stepper._halo_mode = mode
        stepper._peer_state = {}
        stepper._peer_barrier_pending = False
    return mode


def _setup_peer_buffers(stepper, key, f, h):
    """Move ``f`` into a symmetric buffer, create the symmetric twin, exchange the
    geometry of the x neighbours."""
    pg = stepper._peer
    decomp = stepper.decomp
    A, ptrsA = pg.symmetric_like(f)
    B, ptrsB = pg.symmetric_like(f)
    A.copy_(f)
    B.copy_(f)
    f.data = A
    stepper._twins[key] = B
    stepper.__dict__.setdefault("_fresh", {}).pop(key, None)
    shapes = pg.gather_shapes(f.shape)
    lo = decomp.rankID(decomp.rx - 1, decomp.ry, decomp.rz)
    hi = decomp.rankID(decomp.rx + 1, decomp.ry, decomp.rz)
    vol = [int(np.prod(sh[-3:])) for sh in shapes]
    stepper._peer_state[key] = dict(
        ids={A.data_ptr(): 0, B.data_ptr(): 1}, bufs=(A.view(A.shape), B.view(B.shape)),
        ptrs=[ptrsA, ptrsB], lo=lo, hi=hi,
        lo_Nx=shapes[lo][-3] - 2 * h, lo_stride=vol[lo], hi_stride=vol[hi])


def _adopt_into_peer_buffers(stepper, key, f):
    """A tensor that does not live in the symmetric buffers was passed (e.g. freshly
    uploaded initial conditions): move it in; the ping-pong restarts from buffer A."""
    ps = stepper._peer_state[key]
    A, B = ps["bufs"]
    if tuple(A.shape) != tuple(f.shape) or A.dtype != f.dtype:
        raise ValueError("array shape/dtype changed between calls of a multi-rank fused "
                         "stepper; construct a new stepper")
    A.copy_(f)
    f.data = A
    stepper._twins[key] = B.view(B.shape)   # a fresh object: .data swaps must not
    stepper.__dict__.setdefault("_fresh", {}).pop(key, None)   # touch the views in ps["bufs"]


def _marker_rows(stepper, args, peer):
    """CTAs of the launch when it carries no energy partials but needs a stage tail (the
    inter-GPU barrier of a multi-rank stepper without energy sums): marker rows."""
    if stepper.fused_energy or peer is None:
        return 0
    f = args["f_in"]
    key = ("rows", tuple(f.shape), f.dtype, args["xchunk"], args["variant"])
    cache = stepper.__dict__.setdefault("_energy_layout", {})
    if key not in cache:
        probe = {k: v for k, v in args.items() if k != "generated"}
        cache[key] = _kernels.scalar_stage(count_slots_only=True, **probe)
    return cache[key]


def fused_scalar_stage(stepper, system, f, common, queue):
    h = common["h"]
    decomp = stepper.decomp
    multi = decomp is not None and decomp.nranks > 1
    if multi and (decomp.proc_shape[1] != 1):
        raise NotImplementedError(
            "the fused stage supports slab decompositions proc_shape=(P, 1, 1); use the "
            "element-wise stepper with FiniteDifferencer for (Px, Py, 1)")
    if decomp is not None and tuple(decomp.halo_shape) != (h, h, h):
        raise ValueError("decomp.halo_shape differs from the stepper's halo_shape")

    key = system.f_name
    mode = _halo_mode(stepper, decomp, multi)
    if mode == "peer":
        if key not in stepper._peer_state:
            _setup_peer_buffers(stepper, key, f, h)
        elif f.data_ptr() not in stepper._peer_state[key]["ids"]:
            _adopt_into_peer_buffers(stepper, key, f)
    twin = stepper._twins.get(key)
    if twin is None or twin.shape != f.shape or twin.dtype != f.dtype \
            or twin.device != f.device:
        twin = torch.empty_like(f)
        # halo edges/corners of the twin are never written by the kernel; give them
        # defined values once
        twin.copy_(f)
        stepper._twins[key] = twin
        stepper.__dict__.setdefault("_fresh", {}).pop(key, None)

    if not halos_are_fresh(stepper, key, f):
        # halos of the user's array are of unknown age: fill them
        if mode == "peer" and stepper._peer_barrier_pending:
            stepper._peer.barrier_kernel(queue, stepper.strict_fp, stepper.launch_flags)
            stepper._peer_barrier_pending = False
        fill_halos(stepper, f, h, queue)

    Nx = f.shape[-3] - 2 * h
    args = dict(common)
    args.update(fused=True, f_in=f, f_out=twin, dx=stepper.dx)
    tune = stepper.tune or {}
    args.update(xchunk=tune.get("xchunk", 0), variant=tune.get("variant", 0),
                prefetch=tune.get("prefetch", 0), launch_flags=stepper.launch_flags)
    if stepper.__dict__.get("_pws_sched") is None:
        # scheduler words of the persistent kernel variant (zero between launches)
        stepper._pws_sched = torch.zeros(2, dtype=torch.int64, device=f.device)
    args.update(pws_sched=stepper._pws_sched)

    nslots = 0
    if stepper.fused_energy:
        if stepper._V_values is None:
            raise NotImplementedError(
                "fused energy sums need the potential V: construct the stepper from a "
                "ScalarSector")
        args.update(V_terms=stepper._V_values)
        ranges = _launch_ranges(Nx, h, multi and mode != "peer", args["xchunk"] or 16)
        lkey = (tuple(f.shape), f.dtype, multi, mode, args["xchunk"], args["variant"])
        cache = stepper.__dict__.setdefault("_energy_layout", {})
        if lkey not in cache:
            probe = {k: v for k, v in args.items() if k != "generated"}
            cache[lkey] = [_kernels.scalar_stage(x_range=r, count_slots_only=True, **probe)
                           for r in ranges]
        counts = cache[lkey]
        total = sum(counts)
        if stepper.energy_partials is None or stepper.energy_partials.shape[0] < total:
            # rows start out (and are left by the in-kernel tail) as "not yet published"
            stepper.energy_partials = _kernels.sentinel_rows(total, f.device)
        stepper.energy_nslots = total
        stepper.energy_nvals = 2 * system.nscalars + 1
        offsets = [sum(counts[:i]) for i in range(len(counts))]
        energy_kw = [dict(energy_partials=stepper.energy_partials, energy_slot_offset=o)
                     for o in offsets]
        nslots = total
    else:
        energy_kw = [{}, {}]
    stepper._sums_folded = False

    if not multi:
        # one launch; its last CTA folds the energy sums and, with an attached
        # DeviceExpansion, advances the background: no other kernel in the stage
        tail = stepper._make_tail(fold=stepper.fused_energy, dt=common["dt"],
                                  nrows=_marker_rows(stepper, args, None),
                                  device=f.device)
        _kernels.scalar_stage(wrap=(True, True, True), tail=tail, **args, **energy_kw[0])
        stepper._tail_issued(tail)
    elif mode == "peer":
        ps = stepper._peer_state[key]
        if stepper._peer_barrier_pending:
            # the previous stage's last kernel did not carry the barrier
            stepper._peer.barrier_kernel(queue, stepper.strict_fp, stepper.launch_flags)
            stepper._peer_barrier_pending = False
        out_id = ps["ids"][twin.data_ptr()]
        peer = dict(lo=ps["ptrs"][out_id][ps["lo"]], hi=ps["ptrs"][out_id][ps["hi"]],
                    lo_Nx=ps["lo_Nx"], lo_stride=ps["lo_stride"], hi_stride=ps["hi_stride"])
        stepper._peer_barrier_pending = True
        tail = stepper._make_tail(fold=stepper.fused_energy, dt=common["dt"],
                                  nrows=_marker_rows(stepper, args, peer), device=f.device)
        _kernels.scalar_stage(wrap=(False, True, True), peer=peer, tail=tail, **args,
                              **energy_kw[0])
        stepper._tail_issued(tail)
    else:
        st = _ensure_streams(stepper)
        cur = queue if queue is not None else torch.cuda.current_stream()
        _mark("stage_begin", cur)
        with torch.cuda.stream(cur):
            # boundary slabs first
            ranges = _launch_ranges(Nx, h, True, args["xchunk"] or 16)
            _kernels.scalar_stage(wrap=(False, True, True), x_range=ranges[0], **args,
                                  **energy_kw[0])
            interior = ranges[1] if len(ranges) > 1 else None
            st["boundary_done"].record(cur)
            _mark("boundary_done", cur)
        with torch.cuda.stream(st["comm"]):
            st["comm"].wait_event(st["boundary_done"])
            for req in exchange_x_faces(decomp, twin, h, f.dim() == 4):
                req.wait()
            st["comm_done"].record(st["comm"])
            _mark("comm_done", st["comm"])
        if interior is not None:
            with torch.cuda.stream(cur):
                _kernels.scalar_stage(wrap=(False, True, True), x_range=interior, **args,
                                      **energy_kw[1])
            _mark("interior_done", cur)
        cur.wait_event(st["comm_done"])
        _mark("joined", cur)

    _swap_storage(f, twin)
    mark_halos_fresh(stepper, key, f)
    stepper.__dict__.setdefault("_stage_serial", {})[key] = stepper.__dict__.get("_call_serial")
