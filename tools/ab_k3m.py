"""A/B of the copy-mode reward kernels on rows with 9..32 candidates: CUDA-core row sweep (K3) vs the
tensor-core kernel (K3m) with 2 / 3 bf16 terms and 1 / 2 / 4 trajectory tiles per thread.

Usage (GPU box): python tools/ab_k3m.py [cfg] [batch]
Prints the reward-kernel phase time (CUDA events inside the library) and the max relative deviation of the
residuals from K3's; the K3m context is rebuilt per split (the split is fixed when the records are built)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
from gflownet_spai_b200 import synth
from gflownet_spai_b200.env import SpaiContext


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    pb = synth.make_problem(cfg, 1.0)
    coo = pb.a.tocoo()
    dev = torch.device("cuda", 0)
    acts, lens = bench.device_trajectories(pb.num_edges, batch, 0, dev, 0.5)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    for split in ("off", "3", "2"):
        os.environ.pop("SPAI_K3_MMA", None)
        if split == "off":
            os.environ["SPAI_K3_MMA"] = "0"
        else:
            os.environ["SPAI_K3M_SPLIT"] = split
        ctx = SpaiContext(pb.n, pb.edge_row, pb.edge_col, pb.edge_val, coo.row, coo.col, coo.data, device=0)
        ctx.enable_timing(True)
        for ntm in (("0",) if split == "off" else ("4", "2", "1")):
            os.environ["SPAI_K3M_NTM"] = ntm if ntm != "0" else "8"
            ms = []
            for it in range(4):
                flush.zero_()
                out = ctx.reward_batch(acts, 0.5, "copy", torch.float32, lengths=lens)
                torch.cuda.synchronize()
                if it:
                    ms.append(ctx.last_timing().ms_reward)
            res = out["residual"].cpu().numpy()
            if ref is None:
                ref = res
            row = {"cfg": cfg, "B": batch, "kernel": "k3_copy_kernel" if split == "off" else f"k3m split={split} ntm<={ntm}",
                   "reward_ms": float(np.median(ms)), "step_ms": ctx.last_timing().ms_total,
                   "max_rel_dev_vs_k3": float(np.max(np.abs(res - ref) / ref)),
                   "row_evals_per_s": batch * pb.n / (float(np.median(ms)) / 1e3)}
            print(json.dumps(row), flush=True)
        ctx.close()
        del ctx
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
