"""Probe (2+ GPUs): does torch's symmetric memory (peer-mapped buffers + device-side barrier) work on this box, and how
fast is a copy-engine pull of a peer buffer?  torchrun --nproc-per-node N tools/symm_probe.py"""
import os, time
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import torch.distributed._symmetric_memory as symm
try:
    rows, cols = 43872, 512
    t = symm.empty((2, rows, cols), dtype=torch.bfloat16, device=torch.device("cuda", lr))
    hdl = symm.rendezvous(t, dist.group.WORLD)
    t.fill_(rank + 1)
    hdl.barrier(channel=0)
    dst = torch.empty((world, rows, cols), dtype=torch.bfloat16, device="cuda")
    side = torch.cuda.Stream()
    peers = [j for j in range(world) if j != rank]
    bufs = {j: hdl.get_buffer(j, (rows, cols), torch.bfloat16, 0) for j in peers}
    torch.cuda.synchronize()
    for it in range(3):
        hdl.barrier(channel=it % 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in peers:
            dst[j].copy_(bufs[j])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    ok = all(float(dst[j].float().mean()) == j + 1 for j in peers)
    gb = len(peers) * rows * cols * 2 / 1e9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(20):
        hdl.barrier(channel=0)
    e1.record(); torch.cuda.synchronize()
    print(f"rank {rank}: symmetric memory OK={ok}; pulled {gb * 1e3:.0f} MB from {len(peers)} peer(s) in {ms:.3f} ms = {gb / ms * 1e3:.0f} GB/s; "
          f"barrier {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
except Exception as e:
    print(f"rank {rank}: symmetric memory FAILED: {type(e).__name__}: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()
