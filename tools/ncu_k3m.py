"""One K3m launch (and one K3 launch) on cfg3 for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:'k3m_kernel|k3_copy_kernel' -c 2 -o gpurun_out/r2_k3m python tools/ncu_k3m.py [cfg] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
from gflownet_spai_b200 import synth
from gflownet_spai_b200.env import SpaiContext


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    dev = torch.device("cuda", 0)
    p = synth.make_problem(cfg)
    coo = p.a.tocoo()
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=0)
    acts, lens = bench.device_trajectories(p.num_edges, batch, 0, dev, 0.5)
    ctx.reward_batch(acts, 0.5, "copy", torch.float32, lengths=lens)
    os.environ["SPAI_K3_MMA"] = "0"
    ctx.reward_batch(acts, 0.5, "copy", torch.float32, lengths=lens)
    torch.cuda.synchronize()
    ctx.close()


if __name__ == "__main__":
    main()
