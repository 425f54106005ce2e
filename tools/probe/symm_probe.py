"""Probe: does torch symmetric memory (NVLS multicast) work on this box?  Run under torchrun."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm_mem.empty(1024 * 1024, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:3], "multicast_ptr", hex(hdl.multicast_ptr),
          "signal_pad_ptrs", len(hdl.signal_pad_ptrs), flush=True)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer value", float(peer[0]), flush=True)
    hdl.barrier()
except Exception as e:
    import traceback
    traceback.print_exc()
dist.destroy_process_group()
