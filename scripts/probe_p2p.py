"""Probe: peer-memory mechanisms available between ranks of one node (run under torchrun)."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
peer = (rank + 1) % world
print(rank, "can_access_peer", torch.cuda.can_device_access_peer(rank, peer), flush=True)

# --- 1. symmetric memory
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((2, 516, 516), dtype=torch.float64, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(float(rank))
    hdl.barrier()
    pbuf = hdl.get_buffer(peer, (2, 516, 516), torch.float64)
    src = torch.full((2, 516, 516), 100.0 + rank, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for _ in range(3):
        pbuf.copy_(src)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pbuf.copy_(src)
    e1.record(); torch.cuda.synchronize()
    hdl.barrier()
    ok = float(t[0, 0, 0]) == 100.0 + (rank - 1) % world
    print(rank, "symm_mem OK" if ok else "symm_mem WRONG", "copy 4.26MB: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), flush=True)
    e0.record()
    for _ in range(20):
        hdl.barrier()
    e1.record(); torch.cuda.synchronize()
    print(rank, "symm barrier: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), flush=True)
except Exception as exc:
    print(rank, "symm_mem FAILED", repr(exc)[:300], flush=True)

# --- 2. NCCL send/recv latency for the same size
a = torch.zeros((2, 516, 516), dtype=torch.float64, device=dev); b = torch.zeros_like(a)
def xchg():
    ops = [dist.P2POp(dist.isend, a, peer), dist.P2POp(dist.irecv, b, (rank - 1) % world)]
    for r in dist.batch_isend_irecv(ops): r.wait()
for _ in range(5): xchg()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): xchg()
e1.record(); torch.cuda.synchronize()
print(rank, "nccl send/recv 4.26MB: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), flush=True)
s = torch.zeros(16, dtype=torch.float64, device=dev)
for _ in range(5): dist.all_reduce(s)
torch.cuda.synchronize()
e0.record()
for _ in range(50): dist.all_reduce(s)
e1.record(); torch.cuda.synchronize()
print(rank, "nccl allreduce 16 doubles: %.1f us" % (e0.elapsed_time(e1) / 50 * 1e3), flush=True)
dist.barrier(); os._exit(0)
