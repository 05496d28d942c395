import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty(1024, dtype=torch.float64, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", len(hdl.signal_pad_ptrs), flush=True)
t.fill_(rank + 1)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float64)
print(rank, "peer value", peer[0].item(), flush=True)
# write into the peer
peer[rank] = 100 + rank
hdl.barrier()
print(rank, "after", t[:world].tolist(), flush=True)
dist.destroy_process_group()
