#!/usr/bin/env python3
"""2-GPU probe: can ranks map each other's device memory?  (a) torch symmetric memory, (b) raw CUDA IPC
through libjwave_cuda.so (jwc_ipc_export / jwc_ipc_open).  torchrun --nproc-per-node 2 tools/p2p_probe.py"""
import ctypes as C
import os
import sys
import traceback

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
nxt = (rank + 1) % world

try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.float64, device=f"cuda:{local}")
    t.fill_(-1.0)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    hdl.barrier()
    peer = hdl.get_buffer(nxt, t.shape, t.dtype)
    peer.fill_(float(rank + 1))
    torch.cuda.synchronize()
    hdl.barrier()
    print(f"[symm] rank {rank}: my buffer now holds {t[:2].tolist()} (expected {float((rank - 1) % world + 1)})", flush=True)
except Exception:
    print(f"[symm] rank {rank} FAILED", flush=True)
    traceback.print_exc()

try:
    import jwave_b200 as jw
    from jwave_b200 import _lib
    L = _lib.load()
    ctx = jw.CudaContext(local)
    n = 1 << 20
    ptr = C.c_void_p()
    ctx.check(L.jwc_dev_alloc(ctx.handle, n * 8, C.byref(ptr)), "alloc")
    h = (C.c_ubyte * 64)()
    ctx.check(L.jwc_ipc_export(ctx.handle, ptr, h), "export")
    handles = [None] * world
    dist.all_gather_object(handles, bytes(h))
    peer = C.c_void_p()
    hb = (C.c_ubyte * 64).from_buffer_copy(handles[nxt])
    ctx.check(L.jwc_ipc_open(ctx.handle, hb, C.byref(peer)), "open")
    src = torch.full((n,), float(rank + 1), dtype=torch.float64, device=f"cuda:{local}")
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.check(L.jwc_reset_stream(ctx.handle), "stream")
    big = 1 << 27  # 1 GiB for a bandwidth figure
    bptr, bpeer = C.c_void_p(), C.c_void_p()
    ctx.check(L.jwc_dev_alloc(ctx.handle, big * 8, C.byref(bptr)), "alloc big")
    hb2 = (C.c_ubyte * 64)()
    ctx.check(L.jwc_ipc_export(ctx.handle, bptr, hb2), "export big")
    hs = [None] * world
    dist.all_gather_object(hs, bytes(hb2))
    ctx.check(L.jwc_ipc_open(ctx.handle, (C.c_ubyte * 64).from_buffer_copy(hs[nxt]), C.byref(bpeer)), "open big")
    cudart = torch.cuda.cudart()
    torch.cuda.synchronize()
    dist.barrier()
    # copy my src into the peer's small buffer, and time a 1 GiB peer write
    assert int(cudart.cudaMemcpy(peer.value, src.data_ptr(), n * 8, 3)) == 0
    ev0.record()
    for _ in range(3):
        assert int(cudart.cudaMemcpyAsync(bpeer.value, bptr.value, big * 8, 3, torch.cuda.current_stream().cuda_stream)) == 0
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    mine = torch.empty(2, dtype=torch.float64, device=f"cuda:{local}")
    assert int(cudart.cudaMemcpy(mine.data_ptr(), ptr.value, 16, 3)) == 0
    print(f"[ipc] rank {rank}: my buffer now holds {mine.tolist()} (expected {float((rank - 1) % world + 1)}); "
          f"peer write {3 * big * 8 / ev0.elapsed_time(ev1) / 1e6:.0f} GB/s", flush=True)
except Exception:
    print(f"[ipc] rank {rank} FAILED", flush=True)
    traceback.print_exc()

dist.barrier()
dist.destroy_process_group()
