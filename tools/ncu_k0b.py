"""One dense-trajectory reward call on a large pattern, for an ncu capture of the K0b kernels:
    ncu --set full --clock-control none --import-source on -k regex:k0b_ -c 2 -o out python tools/ncu_k0b.py cfg5 128
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
from gflownet_spai_b200 import synth
from gflownet_spai_b200.env import SpaiContext


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = torch.device("cuda", 0)
    p = synth.make_problem(cfg)
    coo = p.a.tocoo()
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=0)
    acts, lens = bench.device_trajectories(p.num_edges, batch, 0, dev, 0.5)
    ctx.kept_mask_words(acts, lengths=lens) if hasattr(ctx, "kept_mask_words") else ctx.reward_batch(acts, 0.5, "copy", torch.float32, lengths=lens)
    torch.cuda.synchronize()
    ctx.close()


if __name__ == "__main__":
    main()
