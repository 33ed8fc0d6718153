"""2-GPU probe: cost of the per-step reward all-gather next to a fixed GPU workload."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.randn(4096, dtype=torch.float64, device=dev); out = torch.empty(world * 4096, dtype=torch.float64, device=dev)
work = torch.randn(8192, 8192, device=dev)
def busy():
    for _ in range(3): torch.mm(work, work)
def timed(fn, n=30):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * (time.perf_counter() - w0) / n
a = timed(lambda: dist.all_gather_into_tensor(out, x))
b = timed(busy)
c = timed(lambda: (busy(), dist.all_gather_into_tensor(out, x)))
if rank == 0: print(f"allgather alone {a}, busy {b}, busy+allgather {c}")
dist.destroy_process_group()
