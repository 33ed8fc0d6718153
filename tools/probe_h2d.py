"""Platform ceiling of concurrent host -> device copies (plain cudaMemcpyAsync from pinned memory, no kernel of this
repository involved): one GPU, pairs, four and all GPUs at once. Explains the end-to-end (`e2e`) scaling of bench.py:
the reference API hands the library `int64[B, T]` action lists in HOST memory, so N ranks pull N x 4.35 GB per step
through whatever the box shares between its GPUs (PCIe switch uplinks, root ports, host DRAM).

    python tools/probe_h2d.py > gpurun_out/h2d_probe.json      # on a multi-GPU box
"""
import json
import time

import torch


def run(devs, bufs_h, bufs_d, streams, reps):
    for d in devs:
        torch.cuda.synchronize(d)
    ev = {}
    t0 = time.perf_counter()
    for d in devs:
        with torch.cuda.device(d), torch.cuda.stream(streams[d]):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                bufs_d[d].copy_(bufs_h[d], non_blocking=True)
            b.record()
            ev[d] = (a, b)
    for d in devs:
        torch.cuda.synchronize(d)
    wall = time.perf_counter() - t0
    nbytes = bufs_h[devs[0]].numel() * reps
    per = {d: nbytes / (ev[d][0].elapsed_time(ev[d][1]) * 1e-3) / 1e9 for d in devs}
    return {"gpus": list(devs), "aggregate_GBps_wall": len(devs) * nbytes / wall / 1e9,
            "per_gpu_GBps_device_events": {str(d): round(v, 2) for d, v in per.items()}}


def main():
    n = torch.cuda.device_count()
    size = 1 << 30
    reps = 4
    bufs_h = {d: torch.empty(size, dtype=torch.uint8).pin_memory() for d in range(n)}
    bufs_d = {d: torch.empty(size, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)}
    streams = {d: torch.cuda.Stream(device=d) for d in range(n)}
    sets = [(0,)]
    if n >= 2:
        sets += [(0, 1)]
    if n >= 4:
        sets += [(0, 2), (0, 1, 2, 3)]
    if n >= 8:
        sets += [(0, 4), (0, 2, 4, 6), tuple(range(8))]
    out = []
    for s in sets:
        run(s, bufs_h, bufs_d, streams, 1)             # warm-up
        out.append(run(s, bufs_h, bufs_d, streams, reps))
    print(json.dumps({"bytes_per_copy": size, "copies": reps, "results": out}))


if __name__ == "__main__":
    main()
