"""Times the per-iteration exchange in isolation: all_reduce of G*K+K fp64 on the NCCL
process group, with and without a compute kernel in front of it (torchrun, N ranks)."""
import os, time, torch, torch.distributed as dist
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 20000 * 20 + 20
buf = torch.zeros(n, dtype=torch.float64, device="cuda")
big = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
for name, pre in (("allreduce only", False), ("after a ~0.1 ms kernel", True)):
    for _ in range(20):
        dist.all_reduce(buf)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        if pre:
            big.add_(1.0)
        dist.all_reduce(buf)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("%-28s %.1f us per iteration (%d doubles, world %d)" % (name, e0.elapsed_time(e1) * 1000 / 200, n, dist.get_world_size()))
torch.cuda.synchronize(); e0.record()
for _ in range(200):
    big.add_(1.0)
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print("kernel alone                 %.1f us" % (e0.elapsed_time(e1) * 1000 / 200))
dist.destroy_process_group()
