"""Multi-GPU plumbing probe (run under torchrun on the GPU box): which NCCL does a dlopen("libnccl.so.2") reach from
inside a torch process, does torch's symmetric memory give peer + multicast (NVLS) pointers here, and what does an
NCCL all-gather of one band cost.  Diagnostic only; prints JSON lines."""
import ctypes
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
out = {"rank": rank, "world": world}
try:
    lib = ctypes.CDLL("libnccl.so.2")
    v = ctypes.c_int(0)
    lib.ncclGetVersion(ctypes.byref(v))
    out["dlopen_nccl_version"] = v.value
    out["torch_nccl_version"] = list(torch.cuda.nccl.version())
    maps = [l.split()[-1] for l in open("/proc/self/maps") if "libnccl" in l]
    out["nccl_mapped"] = sorted(set(maps))
except Exception as e:  # noqa: BLE001
    out["dlopen_nccl_error"] = repr(e)
try:
    import torch.distributed._symmetric_memory as symm

    t = symm.empty(64 << 20, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD)
    out["symm_buffer_ptrs"] = [hex(p) for p in h.buffer_ptrs]
    out["symm_multicast_ptr"] = hex(h.multicast_ptr) if getattr(h, "multicast_ptr", 0) else 0
    out["symm_signal_pads"] = len(h.signal_pad_ptrs)
    t.fill_(rank + 1)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.uint8)
    out["peer_read"] = int(peer[0].item())
except Exception as e:  # noqa: BLE001
    out["symm_error"] = repr(e)[:400]
# NCCL all-gather of equal bands, in place
for mb in (4, 16):
    n = mb << 20
    buf = torch.zeros(n * world, dtype=torch.uint8, device=dev)
    band = buf[rank * n:(rank + 1) * n]
    for _ in range(5):
        dist.all_gather_into_tensor(buf, band)
    torch.cuda.synchronize()
    dist.barrier()
    evs = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.all_gather_into_tensor(buf, band)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    out[f"allgather_{mb}MB_per_rank_ms_median"] = round(ts[len(ts) // 2], 4)
    out[f"allgather_{mb}MB_per_rank_ms_min"] = round(ts[0], 4)
print(json.dumps(out), flush=True)
dist.destroy_process_group()
