"""A/B of the two copy-kernel (K3) variants on one problem, in one process.

Usage (GPU box): python tools/ab_k3.py [cfg] [batch] [delete_fraction]
Builds the problem once, draws `batch` trajectories that delete about
`delete_fraction` of the candidate edges each (ids drawn with replacement:
duplicates are legal input), and times reward_batch with the kernel forced
through SPAI_K3_COMPACT=0/1 / SPAI_K3_SPARSE=0/1 and with the host's own choice.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from gflownet_spai_b200 import synth
from gflownet_spai_b200.env import SpaiContext


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
    t0 = time.time()
    pb = synth.make_problem(cfg, 1.0)
    coo = pb.a.tocoo()
    ctx = SpaiContext(pb.n, pb.edge_row, pb.edge_col, pb.edge_val, coo.row, coo.col, coo.data, device=0)
    e = pb.num_edges
    print(f"{cfg}: n={pb.n} E={e} built in {time.time() - t0:.1f}s", flush=True)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    t = max(1, int(e * frac))
    acts = torch.randint(0, e, (batch, t), generator=g, device="cuda", dtype=torch.int64)
    ctx.enable_timing(True)
    ref = None
    for label, env, sparse in (("dense-variant", "0", "0"), ("compact-variant", "1", "0"),
                               ("deletion-driven (K3s)", None, "1"), ("auto", None, None)):
        for key, val in (("SPAI_K3_COMPACT", env), ("SPAI_K3_SPARSE", sparse)):
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val
        for _ in range(2):
            out = ctx.reward_batch(acts, 0.5, mode="copy", dtype=torch.float32)
        torch.cuda.synchronize()
        ks = []
        for _ in range(5):
            out = ctx.reward_batch(acts, 0.5, mode="copy", dtype=torch.float32)
            torch.cuda.synchronize()
            ks.append(ctx.last_timing())
        rw = out["reward"].cpu().numpy()
        if ref is None:
            ref = rw
        same = float(np.max(np.abs(rw - ref) / np.maximum(1e-300, np.abs(ref))))
        names = ("ms_masks", "ms_transpose", "ms_reward", "ms_finalize", "ms_total")
        print(label, {k: round(float(np.median([getattr(x, k) for x in ks])), 3) for k in names},
              "max rel diff vs dense-variant:", same, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
