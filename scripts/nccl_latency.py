"""diagnostic: latency of small NCCL collectives / point-to-point on this box (torchrun, one rank per GPU)"""
import os, sys, time
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    print("can_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1), flush=True)
t = torch.ones(20, dtype=torch.float64, device=dev)
big = torch.ones(2048, 4, dtype=torch.float64, device=dev)
buf = torch.empty_like(big)
for name, fn in [("all_reduce 160 B", lambda: dist.all_reduce(t)),
                 ("all_gather 32 B", lambda: dist.all_gather([torch.empty(4, dtype=torch.float64, device=dev) for _ in range(world)], t[:4])),
                 ("sendrecv 64 KB", lambda: [r.wait() for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, big, (rank + 1) % world), dist.P2POp(dist.irecv, buf, (rank - 1) % world)])])]:
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 200
    # with a host sync after every op (what a control loop that reads a result pays)
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
        torch.cuda.synchronize()
    dts = (time.perf_counter() - t0) / 200
    if rank == 0:
        print("%-18s pipelined %.1f us   with host sync %.1f us" % (name, dt * 1e6, dts * 1e6), flush=True)
dist.destroy_process_group()
