"""
Peer memory between the ranks of one node (NVLink / NVSwitch): symmetric buffers whose
peer-mapped addresses are handed to the kernels, so that

* the fused stage kernel stores its boundary planes straight into the neighbours'
  x halos (``psb_scalar_stage_fused``, ``peer_*`` arguments) -- compute and halo
  exchange in one kernel, no pack, no send/recv, no side stream;
* ``psb_expansion_stage_peer`` all-gathers the energy sums, acts as the inter-GPU
  barrier between stages and advances the FLRW background in one tiny kernel.

Replaces, for x-slab decompositions on one node, the reference's host-staged
``MPI.Sendrecv`` exchange (``pystella/decomp.py:402-419``) and pickled
``allreduce`` (``decomp.py:470-491``).  Uses ``torch.distributed._symmetric_memory``
for allocation / address exchange only.
"""

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from pystella_b200 import _lib

__all__ = ["PeerGroup", "peer_memory_available"]


def peer_memory_available(decomp):
    if decomp is None or decomp.nranks < 2 or decomp.comm is None:
        return False
    if dist.get_backend(decomp.comm) != "nccl":
        return False
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
    except Exception:  # noqa: BLE001
        return False
    return True


class PeerGroup:
    """Symmetric allocations + peer pointer tables of one DomainDecomposition."""

    def __init__(self, decomp, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm = symm_mem
        self.decomp = decomp
        self.device = device
        self.group_name = decomp.comm.group_name
        P = decomp.nranks
        self.P = P
        # [2][P][16] double slots followed by [2][P] uint64 flags
        self.nslot = 2 * P * _lib.ENERGY_STRIDE
        n = self.nslot + 2 * P
        self.sync_buf = symm_mem.empty(n, dtype=torch.float64, device=device)
        self.sync_buf.zero_()
        self.sync_hdl = symm_mem.rendezvous(self.sync_buf, self.group_name)
        self.sync_hdl.barrier()
        slot_ptrs, flag_ptrs = [], []
        for q in range(P):
            base = self.sync_hdl.get_buffer(q, (n,), torch.float64).data_ptr()
            slot_ptrs.append(base)
            flag_ptrs.append(base + self.nslot * 8)
        self._slot_tab = torch.tensor(slot_ptrs, dtype=torch.int64, device=device)
        self._flag_tab = torch.tensor(flag_ptrs, dtype=torch.int64, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        self.args = _lib.PeerArgs()
        self.args.nranks = P
        self.args.rank = decomp.rank
        self.args.peer_slots = self._slot_tab.data_ptr()
        self.args.peer_flags = self._flag_tab.data_ptr()
        self.args.epoch = self.epoch.data_ptr()
        self._dummy = torch.zeros(32, dtype=torch.float64, device=device)
        self._keep = []

    def symmetric_like(self, t):
        """
        A symmetric buffer viewed with ``t``'s shape/dtype (sized for the largest rank so
        uneven slabs work) and the list of its peer-mapped base addresses, one per rank.
        """
        numel = int(self.decomp.allreduce(int(t.numel()), op="max"))
        buf = self.symm.empty(numel, dtype=t.dtype, device=self.device)
        hdl = self.symm.rendezvous(buf, self.group_name)
        ptrs = [hdl.get_buffer(q, (numel,), t.dtype).data_ptr() for q in range(self.P)]
        self._keep.append((buf, hdl))
        return buf[:t.numel()].view(t.shape), ptrs

    def gather_shapes(self, shape):
        """Every rank's padded field shape (list indexed by rank)."""
        out = [None] * self.P
        dist.all_gather_object(out, tuple(int(x) for x in shape), group=self.decomp.comm)
        return out

    def barrier_kernel(self, queue=None, strict=False, launch_flags=0):
        """Inter-GPU barrier as a kernel on ``queue`` (orders peer stores)."""
        from pystella_b200 import _kernels
        _kernels.expansion_stage_peer(self.args, self._dummy, 0, 1.0, 1.0, 0.0, 0.0, 0.0,
                                      self._dummy, self._dummy, queue, strict, launch_flags)
