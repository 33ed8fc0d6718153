"""Probe: how much do K0 (HBM-bound) and K3 (issue-bound) overlap when two half-batches run
on two streams with a stagger? Decides whether a pipelined driver is worth building."""
import sys, torch
sys.path.insert(0, ".")
from gflownet_spai_b200 import synth
from gflownet_spai_b200.env import SpaiContext
from bench import device_trajectories

p = synth.make_problem("cfg2"); coo = p.a.tocoo()
mk = lambda: SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data)
ctxs = [mk() for _ in range(4)]
dev = torch.device("cuda", 0)
acts, lens = device_trajectories(p.num_edges, 4096, 0, dev, 0.5)
streams = [torch.cuda.Stream() for _ in range(4)]

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def single():
    ctxs[0].reward_batch(acts, 0.5, "copy", torch.float32, lengths=lens)

def split(nchunk, stagger_cycles):
    def run():
        cur = torch.cuda.current_stream()
        start = torch.cuda.Event(); start.record(cur)
        step = 4096 // nchunk
        for i in range(nchunk):
            s = streams[i % len(streams)]
            s.wait_event(start)
            with torch.cuda.stream(s):
                if stagger_cycles and i: torch.cuda._sleep(int(stagger_cycles * i))
                ctxs[i % len(ctxs)].reward_batch(acts[i * step:(i + 1) * step], 0.5, "copy", torch.float32, lengths=lens[i * step:(i + 1) * step])
        for s in streams[:nchunk]:
            cur.wait_stream(s)
    return run

print("single call           ", timed(single))
for nchunk in (2, 4):
    for stag_ms in (0.0, 0.2, 0.4):
        print(f"{nchunk} chunks, stagger {stag_ms} ms", timed(split(nchunk, stag_ms * 1.9e6)))
