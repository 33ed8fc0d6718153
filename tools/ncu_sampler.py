"""One launch of every sampler kernel for an `ncu --set full` capture (never a source of bench numbers):

    ncu --set full --clock-control none --import-source on -k regex:'k4g_|k4p_|k4_sample' -c 12 -o gpurun_out/r2_sampler python tools/ncu_sampler.py

cfg2's action count (A = 524 281), B = 592 samples (4 per SM): K4g count + order, K4p 4096 steps, K4 one step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from gflownet_spai_b200 import synth
from gflownet_spai_b200.env import SpaiContext


def main():
    bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    dev = torch.device("cuda", 0)
    p = synth.make_problem("cfg1")
    coo = p.a.tocoo()
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=0)
    a = 524281
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    logits = (torch.randn(a, generator=g, device=dev) * 0.25).contiguous()
    taken, length = ctx.sample_taken(logits, bsz, 12345, 0)
    acts = ctx.sample_order(logits, length, 12345, 0)
    words = (a + 31) // 32
    tk = torch.zeros((bsz, words), dtype=torch.int32, device=dev)
    dn = torch.zeros(bsz, dtype=torch.uint8, device=dev)
    ctx.sample_steps(logits, tk, dn, 4096, seed=3)
    act = torch.empty(bsz, dtype=torch.int64, device=dev)
    pr = torch.empty(bsz, dtype=torch.float32, device=dev)
    ctx.sample_step(logits, tk, torch.rand(bsz, device=dev), dn, act, pr)
    torch.cuda.synchronize()
    print("ok", int(length.sum()), acts.shape)
    ctx.close()


if __name__ == "__main__":
    main()
