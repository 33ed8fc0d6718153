"""One launch of every kernel the profiles/ summaries cover, for an `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on -k regex:'k0b_|k1_fill|k4_|k4g_|k4p_|k3s_sparse|k3t_lookup|k0_mask_build|k0_transpose|k3_copy|k3m_kernel' \
        -c 60 -o gpurun_out/r2_targets python tools/ncu_targets.py

Workloads: cfg2 B=4096 headline step (K0 smem with row lengths, transpose, K3t; K1 at context creation),
cfg2 short trajectories (K3s), K4 / K4g / K4p at A = 524 281, cfg3 B=256 dense trajectories (K0b sort + build, K3m on the
tensor cores, then the CUDA-core row sweep K3 with SPAI_K3_MMA=0).
Never a source of bench numbers (ncu serialises and replays).
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
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2,k4,cfg3"
    dev = torch.device("cuda", 0)
    if "cfg2" in which:
        p = synth.make_problem("cfg2")
        coo = p.a.tocoo()
        ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=0)   # K1
        acts, lens = bench.device_trajectories(p.num_edges, 4096, 0, dev, 0.5)
        ctx.reward_batch(acts, 0.5, "copy", torch.float32, lengths=lens)                                     # K0, transpose, K3t
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        short = torch.randint(0, p.num_edges, (4096, 5243), generator=g, device=dev, dtype=torch.int64)      # 1 % deletions -> K3s
        ctx.reward_batch(short, 0.5, "copy", torch.float32)
        if "k4" in which:
            a = p.num_edges + 1
            bsz = 256
            logits = torch.randn(a, device=dev)
            taken = torch.zeros((bsz, (a + 31) // 32), dtype=torch.int32, device=dev)
            done = torch.zeros(bsz, dtype=torch.uint8, device=dev)
            act = torch.empty(bsz, dtype=torch.int64, device=dev)
            prob = torch.empty(bsz, dtype=torch.float32, device=dev)
            for _ in range(2):
                ctx.sample_step(logits, taken, torch.rand(bsz, device=dev), done, act, prob)
            tk2, ln2 = ctx.sample_taken(logits, 592, 12345, 0)                                               # K4g count
            ctx.sample_order(logits, ln2, 12345, 0)                                                          # K4g max / ntable / order
            tk3 = torch.zeros((592, (a + 31) // 32), dtype=torch.int32, device=dev)
            ctx.sample_steps(logits, tk3, torch.zeros(592, dtype=torch.uint8, device=dev), 2048, seed=3)     # K4p
        torch.cuda.synchronize()
        ctx.close()
        del acts, lens, short
    if "cfg3" in which:
        p = synth.make_problem("cfg3")
        coo = p.a.tocoo()
        ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=0)
        acts, lens = bench.device_trajectories(p.num_edges, 256, 0, dev, 0.5)
        ctx.reward_batch(acts, 0.5, "copy", torch.float32, lengths=lens)                                     # K0b sort + build, K3m
        os.environ["SPAI_K3_MMA"] = "0"
        ctx.reward_batch(acts, 0.5, "copy", torch.float32, lengths=lens)                                     # the row sweep K3
        os.environ.pop("SPAI_K3_MMA")
        torch.cuda.synchronize()
        ctx.close()


if __name__ == "__main__":
    main()
