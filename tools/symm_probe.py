"""Developer probe: does torch symmetric memory (peer-mapped buffers over NVLink) work on this box?
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/symm_probe.py"""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(2 * 100008, dtype=torch.float64, device=dev)
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "signal", [hex(p) for p in h.signal_pad_ptrs], "pad bytes", h.signal_pad_size,
      "multicast_ptr", hex(h.multicast_ptr), flush=True)
t.fill_(float(rank + 1))
h.barrier(0)
peer = h.get_buffer((rank + 1) % world, (8,), torch.float64, 0)
print(rank, "peer value", peer[:2].tolist(), flush=True)
h.barrier(0)
# timing: NCCL all-reduce of 800 KB vs barrier
x = torch.randn(100000, dtype=torch.float64, device=dev)
for name, fn in (("nccl_allreduce_800KB", lambda: dist.all_reduce(x)), ("symm_barrier", lambda: h.barrier(0))):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(name, e0.elapsed_time(e1) / 200 * 1e3, "us", flush=True)
dist.destroy_process_group()
