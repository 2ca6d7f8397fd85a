"""
Domain decomposition over ``torch.distributed`` (NCCL over NVLink on device
tensors; any backend for the rank arithmetic).  Same constructor, attributes and
methods as ``pystella/decomp.py:112-725``; what changed underneath:

* the reference stages every face through host memory and blocking
  ``MPI.Sendrecv`` (``decomp.py:402-419, 431-448``); here faces go GPU-to-GPU.
  x faces of a C-ordered padded array are contiguous, so they are sent straight
  out of / received straight into the field tensor (no pack kernel); only y faces
  are packed (``psb_halo_pack`` / ``psb_halo_unpack``);
* non-decomposed axes are wrapped in place by ``psb_halo_wrap`` (the
  ``pack_unpack_*`` kernels, ``decomp.py:181-182, 208-209, 232-233``);
* ``allreduce`` / ``bcast`` are ``torch.distributed`` collectives instead of
  pickled MPI object collectives (``decomp.py:451-491``).

One process per GPU (``torchrun``); with no initialised process group the
decomposition is single-rank, exactly like the reference without mpi4py.
"""

import logging

import numpy as np
import torch
import torch.distributed as dist

from pystella_b200 import _kernels, _lib
from pystella_b200._lib import PsbError

logger = logging.getLogger(__name__)

__all__ = ["DomainDecomposition"]

_REDUCE_OPS = {
    None: dist.ReduceOp.SUM, "sum": dist.ReduceOp.SUM, "avg": dist.ReduceOp.SUM,
    "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "prod": dist.ReduceOp.PRODUCT,
}


class DomainDecomposition:
    """
    :arg proc_shape: 3-tuple, ranks per axis; ``proc_shape[2]`` must be 1
        (``decomp.py:129-130``).
    :arg halo_shape: int or 3-tuple, halo layers per axis.
    :arg rank_shape: optional fixed rank-local shape.
    :arg grid_shape: optional global shape (then ``rank_shape`` is derived with the
        uneven split of ``decomp.py:322-337``); pass only one of the two.

    Attributes: ``comm`` (the process group or *None*), ``rank``, ``nranks``,
    ``proc_shape``, ``rank_shape``, ``halo_shape``, ``rank_tuple``.
    """

    def __init__(self, proc_shape, halo_shape, rank_shape=None, grid_shape=None,
                 group=None):
        self.proc_shape = tuple(proc_shape)
        self.halo_shape = ((halo_shape,) * 3 if isinstance(halo_shape, (int, np.integer))
                           else tuple(halo_shape))
        self.buffer_arrays = {}
        self.rank_shape = tuple(rank_shape) if rank_shape is not None else None
        #: test hook: object with wrap/pack/unpack for host tensors (tests only)
        self.host_ops = None

        if dist.is_available() and dist.is_initialized():
            self.comm = group if group is not None else dist.group.WORLD
            self.rank = dist.get_rank(self.comm)
            self.nranks = dist.get_world_size(self.comm)
        else:
            self.comm = None
            self.rank = 0
            self.nranks = 1

        if self.proc_shape[2] != 1:
            raise NotImplementedError("decomposition in z not yet supported")

        if np.prod(self.proc_shape) != self.nranks:
            raise ValueError(
                f"{self.proc_shape} is an invalid decomposition for "
                f"{self.nranks} ranks")

        # decomp.py:137-139
        P = self.proc_shape
        self.rz = self.rank % P[2]
        self.ry = (self.rank - self.rz) // P[2] % P[1]
        self.rx = (self.rank - self.rz - P[2] * self.ry) // P[1]

        if grid_shape is not None:
            if self.rank_shape is not None:
                raise ValueError("Only pass one of rank_shape or grid_shape.")
            self.rank_shape, _ = self.get_rank_shape_start(grid_shape)

    # -- rank arithmetic ------------------------------------------------------
    def rankID(self, rx, ry, rz):
        """Rank of processor-grid site ``(rx, ry, rz)``, periodic (decomp.py:287-296)."""
        P = self.proc_shape
        return (rz % P[2]) + P[2] * ((ry % P[1]) + P[1] * (rx % P[0]))

    @property
    def rank_tuple(self):
        return (self.rx, self.ry, self.rz)

    def get_rank_shape_start(self, grid_shape, rank_tuple=None):
        """
        ``(rank_shape, rank_start)`` of ``rank_tuple`` (default: this rank) for a
        global array of shape ``grid_shape``; remainders go to the first ranks
        (decomp.py:305-337).
        """
        rank_tuple = rank_tuple or self.rank_tuple
        shape, start = [], []
        for N, P, r in zip(grid_shape, self.proc_shape, rank_tuple):
            q, rem = divmod(N, P)
            shape.append(q + (1 if rem > r else 0))
            start.append(r * q + min(r, rem))
        return tuple(shape), tuple(start)

    # -- collectives ----------------------------------------------------------
    def Barrier(self):
        if self.comm is not None:
            dist.barrier(group=self.comm)

    def _coll_device(self):
        if self.comm is not None and dist.get_backend(self.comm) == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def bcast(self, x, root):
        """Broadcast an arbitrary Python object from ``root`` (decomp.py:451-468)."""
        if self.comm is not None:
            box = [x]
            dist.broadcast_object_list(box, src=root, group=self.comm,
                                       device=self._coll_device())
            x = box[0]
        return x

    def allreduce(self, rank_reduction, op=None):
        """
        Reduce a per-rank number (or small tensor / array) over all ranks
        (decomp.py:470-491).  ``op``: ``"sum"`` (default), ``"max"``, ``"min"``,
        ``"prod"`` or a ``torch.distributed.ReduceOp``.
        """
        if self.comm is None:
            return rank_reduction
        rop = _REDUCE_OPS[op] if (op is None or isinstance(op, str)) else op
        if isinstance(rank_reduction, torch.Tensor):
            t = rank_reduction.to(self._coll_device())
            dist.all_reduce(t, op=rop, group=self.comm)
            return t.to(rank_reduction.device)
        arr = np.asarray(rank_reduction)
        t = torch.as_tensor(arr.astype(np.float64 if arr.dtype.kind == "f"
                                       else np.int64)).reshape(-1)
        t = t.to(self._coll_device())
        dist.all_reduce(t, op=rop, group=self.comm)
        out = t.cpu().numpy().reshape(arr.shape).astype(arr.dtype)
        if arr.ndim == 0:
            return out.item() if not isinstance(rank_reduction, np.ndarray) else out
        return out

    # -- local kernels (device) or injected host ops (tests) -------------------
    def _wrap(self, fx, mask, queue):
        if _lib.on_device(fx):
            _kernels.halo_wrap(fx, self.halo_shape, mask, queue)
        elif self.host_ops is not None:
            self.host_ops.wrap(fx, self.halo_shape, mask)
        else:
            raise PsbError(f"share_halos: array lives on {fx.device}; halo kernels run "
                           "on the GPU only (no CPU fallback)")

    def _pack(self, fx, buf, axis, queue, unpack=False):
        if _lib.on_device(fx):
            _kernels.halo_pack(fx, buf, self.halo_shape, axis, queue, unpack=unpack)
        elif self.host_ops is not None:
            self.host_ops.pack(fx, buf, self.halo_shape, axis, unpack)
        else:
            raise PsbError(f"share_halos: array lives on {fx.device}; halo kernels run "
                           "on the GPU only (no CPU fallback)")

    def _get_or_create_halo_buffers(self, shape, dtype, device):
        key = (tuple(shape), dtype, str(device))
        if key not in self.buffer_arrays:
            self.buffer_arrays[key] = (torch.zeros(shape, dtype=dtype, device=device),
                                       torch.zeros(shape, dtype=dtype, device=device))
        return self.buffer_arrays[key]

    def _exchange(self, sends, recvs, lo_rank, hi_rank):
        """
        ``sends = (to_lo, to_hi)``, ``recvs = (from_hi, from_lo)``: my low face
        goes to the lower neighbour and arrives in *its* high halo, and vice versa
        (decomp.py:405-415).  All four transfers are posted as one batch.
        """
        ops = [
            dist.P2POp(dist.isend, sends[0], lo_rank, group=self.comm),
            dist.P2POp(dist.irecv, recvs[0], hi_rank, group=self.comm),
            dist.P2POp(dist.isend, sends[1], hi_rank, group=self.comm),
            dist.P2POp(dist.irecv, recvs[1], lo_rank, group=self.comm),
        ]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def share_halos(self, queue, fx):
        """
        Fill the halos of the rank-local padded array ``fx`` in place with
        periodic / neighbouring data, axis order z, x, y; every pass spans the
        padded extent of the other axes, so edges and corners are correct
        (decomp.py:365-449).  ``queue`` is a ``torch.cuda.Stream`` or *None*.
        """
        if not isinstance(fx, torch.Tensor):
            raise TypeError("share_halos expects a torch.Tensor")
        if fx.dim() != 3:
            raise ValueError("share_halos expects a 3-D (padded) array")
        if not fx.is_contiguous():
            raise ValueError("share_halos needs a contiguous array")
        h = self.halo_shape
        N = tuple(n - 2 * hi for n, hi in zip(fx.shape, h))
        if self.rank_shape is not None and N != tuple(self.rank_shape) \
                and self.nranks > 1:
            logger.debug("share_halos called with a shape other than rank_shape")

        ctx = torch.cuda.stream(queue) if (queue is not None and fx.is_cuda) else None
        if ctx is not None:
            ctx.__enter__()
        try:
            if h[2] > 0:
                self._wrap(fx, 4, queue)
            if h[0] > 0:
                if self.proc_shape[0] == 1:
                    self._wrap(fx, 1, queue)
                else:
                    lo = self.rankID(self.rx - 1, self.ry, self.rz)
                    hi = self.rankID(self.rx + 1, self.ry, self.rz)
                    self._exchange((fx[h[0]:2 * h[0]], fx[N[0]:N[0] + h[0]]),
                                   (fx[N[0] + h[0]:], fx[:h[0]]), lo, hi)
            if h[1] > 0:
                if self.proc_shape[1] == 1:
                    self._wrap(fx, 2, queue)
                else:
                    shape = (2, N[0] + 2 * h[0], h[1], N[2] + 2 * h[2])
                    send, recv = self._get_or_create_halo_buffers(
                        shape, fx.dtype, fx.device)
                    self._pack(fx, send, 1, queue)
                    lo = self.rankID(self.rx, self.ry - 1, self.rz)
                    hi = self.rankID(self.rx, self.ry + 1, self.rz)
                    self._exchange((send[0], send[1]), (recv[0], recv[1]), lo, hi)
                    self._pack(fx, recv, 1, queue, unpack=True)
        finally:
            if ctx is not None:
                ctx.__exit__(None, None, None)

    # -- padding / gather / scatter ------------------------------------------
    def _interior(self):
        return tuple(slice(hi, -hi) if hi > 0 else slice(None)
                     for hi in self.halo_shape)

    def remove_halos(self, queue, in_array, out_array):
        """Strip the halo padding: ``out = in[h:-h, ...]`` (decomp.py:493-534)."""
        if in_array.dtype != out_array.dtype and not _same_dtype(in_array, out_array):
            raise ValueError("in_array and out_array have different dtypes")
        t_in, t_out = _is_dev(in_array), _is_dev(out_array)
        slc = self._interior()
        if t_in and t_out:
            _kernels.remove_halos(in_array, out_array, self.halo_shape, queue)
        elif t_in:
            out_array[:] = in_array[slc].cpu().numpy()
        elif t_out:
            out_array.copy_(torch.from_numpy(np.ascontiguousarray(in_array[slc])))
        else:
            out_array[:] = in_array[slc]

    def restore_halos(self, queue, in_array, out_array):
        """Add halo padding (halos are *not* filled; decomp.py:598-648)."""
        if not _same_dtype(in_array, out_array):
            raise ValueError("in_array and out_array have different dtypes")
        t_in, t_out = _is_dev(in_array), _is_dev(out_array)
        slc = self._interior()
        if t_in and t_out:
            _kernels.remove_halos(out_array, in_array, self.halo_shape, queue,
                                  restore=True)
        elif t_in:
            out_array[slc] = in_array.cpu().numpy()
        elif t_out:
            out_array[slc] = torch.from_numpy(np.ascontiguousarray(in_array)).to(
                out_array.device)
        else:
            out_array[slc] = in_array

    def gather_array(self, queue, in_array, out_array, root):
        """
        Gather the rank-local (padded) subarrays into the global unpadded
        ``out_array`` on ``root`` (decomp.py:536-596).
        """
        dtype = self.bcast(None if self.rank != root else _np_dtype(out_array), root)
        if _np_dtype(in_array) != dtype:
            raise ValueError("in_array and out_array have different dtypes")
        if self.nranks == 1:
            self.remove_halos(queue, in_array, out_array)
            return
        h = self.halo_shape
        rank_shape = tuple(n - 2 * hi for n, hi in zip(in_array.shape, h))
        full_shape = self.bcast(None if self.rank != root else tuple(out_array.shape),
                                root)
        dev = self._coll_device()
        local = _to_torch(in_array)[self._interior()].contiguous().to(dev)
        if self.rank == root:
            for rx in range(self.proc_shape[0]):
                for ry in range(self.proc_shape[1]):
                    src = self.rankID(rx, ry, 0)
                    shp, start = self.get_rank_shape_start(full_shape, (rx, ry, 0))
                    if src == root:
                        block = local
                    else:
                        block = torch.empty(shp, dtype=local.dtype, device=dev)
                        dist.recv(block, src=src, group=self.comm)
                    slc = tuple(slice(s, s + n) for s, n in zip(start, shp))
                    if isinstance(out_array, torch.Tensor):
                        out_array[slc] = block.to(out_array.device)
                    else:
                        out_array[slc] = block.cpu().numpy()
        else:
            assert tuple(local.shape) == rank_shape
            dist.send(local, dst=root, group=self.comm)

    def scatter_array(self, queue, in_array, out_array, root):
        """
        Scatter the global unpadded ``in_array`` on ``root`` into the interior of
        each rank's padded ``out_array`` (halos untouched; decomp.py:650-725).
        """
        dtype = self.bcast(None if self.rank != root else _np_dtype(in_array), root)
        if _np_dtype(out_array) != dtype:
            raise ValueError("in_array and out_array have different dtypes")
        if self.nranks == 1:
            self.restore_halos(queue, in_array, out_array)
            return
        full_shape = self.bcast(None if self.rank != root else tuple(in_array.shape),
                                root)
        dev = self._coll_device()
        my_shape, _ = self.get_rank_shape_start(full_shape)
        mine = None
        if self.rank == root:
            src_t = _to_torch(in_array)
            for rx in range(self.proc_shape[0]):
                for ry in range(self.proc_shape[1]):
                    dst = self.rankID(rx, ry, 0)
                    shp, start = self.get_rank_shape_start(full_shape, (rx, ry, 0))
                    slc = tuple(slice(s, s + n) for s, n in zip(start, shp))
                    block = src_t[slc].contiguous().to(dev)
                    if dst == root:
                        mine = block
                    else:
                        dist.send(block, dst=dst, group=self.comm)
        else:
            mine = torch.empty(my_shape, dtype=_to_torch_dtype(dtype), device=dev)
            dist.recv(mine, src=root, group=self.comm)
        slc = self._interior()
        if isinstance(out_array, torch.Tensor):
            out_array[slc] = mine.to(out_array.device)
        else:
            out_array[slc] = mine.cpu().numpy()


def _is_dev(x):
    return isinstance(x, torch.Tensor) and _lib.on_device(x)


def _np_dtype(x):
    if isinstance(x, torch.Tensor):
        return np.dtype(str(x.dtype).replace("torch.", ""))
    return np.dtype(x.dtype)


def _to_torch_dtype(dt):
    return getattr(torch, np.dtype(dt).name)


def _same_dtype(a, b):
    return _np_dtype(a) == _np_dtype(b)


def _to_torch(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
