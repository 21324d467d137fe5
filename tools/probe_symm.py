"""Manual 2-GPU probe: does torch.distributed._symmetric_memory work here, and is an NVLS multicast pointer available?"""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else 0,
          "signal_pad", len(h.signal_pad_ptrs), flush=True)
    t.fill_(rank + 1)
    torch.cuda.synchronize(); dist.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.uint8)
    print(rank, "peer sees", peer[:4].tolist(), flush=True)
except Exception as e:
    print(rank, "symm_mem failed:", type(e).__name__, str(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
